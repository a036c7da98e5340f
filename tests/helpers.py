"""Shared case builder for the parity tests: seeded inputs in the reference's test style
(testing/python/operators/test_general_matmul_ops_backend_tl.py:127-226), the oracle output, and the
product-side tensors prepared with the product's own transform."""
from __future__ import annotations

import numpy as np
import torch

import bitblas_oracle as O


def make_case(M, N, K, *, A_dtype="float16", W_dtype="uint4", accum_dtype="float16", out_dtype="float16",
              group_size=-1, with_scaling=False, with_zeros=False, zeros_mode="original", with_bias=False,
              fast_decoding=None, seed=0, int_zeros=True, scale_mag=None):
    g = torch.Generator().manual_seed(seed)
    fmt, bit = O.W_DTYPE_MAP[W_dtype]
    adt = O.TORCH_DTYPE[A_dtype]
    gs = K if group_size in (-1, None) else group_size
    G = K // gs
    if A_dtype == "int8":
        A = torch.randint(-128, 128, (M, K), generator=g, dtype=torch.int8)
    else:
        A = (torch.rand((M, K), generator=g) - 0.5).to(adt)
    fields = torch.randint(0, 2**bit, (N, K), generator=g, dtype=torch.int32)
    scale = zeros = bias = None
    if with_scaling:
        mag = scale_mag if scale_mag is not None else 2.0 / (2**bit)
        scale = (torch.rand((N, G), generator=g) * mag + 0.01).to(adt)
    if with_zeros:
        if zeros_mode == "quantized":
            zq = torch.randint(0, 2**bit, (G, N), generator=g, dtype=torch.int8)
            zeros = torch.from_numpy(O.general_compress(zq.numpy(), bit))
        elif zeros_mode == "original":
            if int_zeros:
                zeros = torch.randint(0, 2**bit, (N, G), generator=g).to(adt)
            else:
                zeros = (torch.rand((N, G), generator=g) * (2**bit - 1)).to(adt)
        else:
            zeros = ((torch.rand((N, G), generator=g) * (2**bit - 1)).to(adt) * scale).to(adt)
    if with_bias:
        bias = torch.rand((N,), generator=g).to(adt) if A_dtype != "int8" else torch.randint(-8, 8, (N,), generator=g, dtype=torch.int8)
    return dict(M=M, N=N, K=K, A=A, fields=fields, scale=scale, zeros=zeros, bias=bias, bit=bit, fmt=fmt,
                cfg=dict(A_dtype=A_dtype, W_dtype=W_dtype, accum_dtype=accum_dtype, out_dtype=out_dtype,
                         group_size=group_size, with_scaling=with_scaling, with_zeros=with_zeros, zeros_mode=zeros_mode,
                         with_bias=with_bias, fast_decoding=fast_decoding))


def oracle_output(case, fast_decoding=True, rows=None):
    c = case["cfg"]
    A = case["A"] if rows is None else case["A"][rows]
    return O.matmul_dequant(A, case["fields"], W_dtype=c["W_dtype"], A_dtype=c["A_dtype"], accum_dtype=c["accum_dtype"],
                            out_dtype=c["out_dtype"], group_size=c["group_size"], with_scaling=c["with_scaling"],
                            with_zeros=c["with_zeros"], zeros_mode=c["zeros_mode"], scale=case["scale"], zeros=case["zeros"],
                            bias=case["bias"], fast_decoding=fast_decoding)


def product_operator(case, M=None):
    import bitblas_b200 as bitblas
    c = case["cfg"]
    cfg = bitblas.MatmulConfig(M=M if M is not None else case["M"], N=case["N"], K=case["K"], A_dtype=c["A_dtype"],
                               W_dtype=c["W_dtype"], accum_dtype=c["accum_dtype"], out_dtype=c["out_dtype"], layout="nt",
                               with_bias=c["with_bias"], group_size=c["group_size"], with_scaling=c["with_scaling"],
                               with_zeros=c["with_zeros"], zeros_mode=c["zeros_mode"], fast_decoding=c["fast_decoding"])
    return bitblas.Matmul(cfg, enable_tuning=False)


def product_weight(op, case, device="cuda"):
    """stored weight from unsigned fields, like the reference tests: matmul.weight_transform(intweight (+maxq))."""
    f8 = case["fields"].to(torch.int8)
    if op.weight_transform is not None:
        return op.weight_transform(f8.cpu()).to(device)
    return f8.to(device)


def run_product(op, case, device="cuda"):
    W = product_weight(op, case, device)
    kw = {}
    if case["scale"] is not None:
        kw["scale"] = case["scale"].to(device)
    if case["zeros"] is not None:
        kw["zeros"] = case["zeros"].to(device)
    if case["bias"] is not None:
        kw["bias"] = case["bias"].to(device)
    out = op(case["A"].to(device), W, **kw)
    torch.cuda.synchronize()
    return out.cpu()


def assert_fp_close(got, ref, what="", max_mismatched_ratio=0.0):
    """north_star tolerance: <= 1e-2 relative for FP accumulate.  Checked two ways: normwise relative error
    <= 1e-2 and the reference's own elementwise criterion (rtol=1e-2, atol=1e-2, torch_assert_close) with NO
    mismatches allowed by default (the reference allows 5 %; bfloat16-output cases with thousands of outputs pass
    max_mismatched_ratio=2e-3: one bf16 ulp is up to 0.78 % of the value, so a double-rounding difference of two ulps
    between the fp32-accumulating kernel and the oracle's A_dtype-accumulate emulation can exceed rtol on single elements)."""
    err = O.rel_fro_error(got, ref)
    assert err <= 1e-2, f"{what}: normwise rel err {err:.3e} > 1e-2"
    O.torch_assert_close(got.float(), ref.float(), rtol=1e-2, atol=1e-2 * max(1.0, float(ref.float().abs().mean())),
                         max_mismatched_ratio=max_mismatched_ratio)
    return err
