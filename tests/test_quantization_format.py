"""CPU: the product's storage-format code (bitblas_b200.quantization + the C++ host routines behind the C ABI)
against the golden vectors from the reference and against the oracle."""
import ctypes
import os

import numpy as np
import pytest
import torch

import bitblas_oracle as O
from bitblas_b200 import _lib
from bitblas_b200.quantization import general_compress, interleave_weight
from bitblas_b200.ops.general_matmul import WeightTransform

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "quant_golden.npz"))


@pytest.mark.parametrize("bits", [4, 2, 1])
def test_python_compress_golden(bits):
    assert np.array_equal(general_compress(G[f"w_b{bits}"], bits), G[f"compress_b{bits}"])


@pytest.mark.parametrize("bits", [4, 2, 1])
@pytest.mark.parametrize("tgt", ["float16", "int8"])
def test_python_interleave_golden(bits, tgt):
    got = interleave_weight(G[f"compress_b{bits}"], bits, tgt)
    assert np.array_equal(got, G[f"ref_cpp_interleave_b{bits}_{tgt}"])
    key = f"interleave_b{bits}_{tgt}"
    if key in G.files:  # cases the reference's numpy function can run
        assert np.array_equal(got, G[key])


@pytest.mark.parametrize("bits", [4, 2, 1])
@pytest.mark.parametrize("tgt", [0, 16, 8])
def test_cpp_host_transform_matches_oracle(bits, tgt):
    rng = np.random.RandomState(bits * 10 + tgt)
    w = torch.from_numpy(rng.randint(0, 2**bits, size=(24, 256)).astype(np.int8))
    got = WeightTransform(bits, tgt)(w)
    exp = O.general_compress(w.numpy(), bits)
    if tgt:
        exp = O.interleave_weight(exp, bits, "int8" if tgt == 8 else "float16")
    assert got.dtype == torch.int8 and tuple(got.shape) == (24, 256 * bits // 8)
    assert np.array_equal(got.numpy(), exp)


def test_cpp_host_error_codes():
    lib = _lib.load()
    buf = (ctypes.c_int8 * 16)()
    assert lib.bb_compress_host(buf, buf, 1, 16, 3) != 0
    assert b"bad arguments" in lib.bb_last_error()
    assert lib.bb_interleave_host(buf, buf, 6, 4, 16) != 0


def test_transform_weight_signed_clamp_and_shift():
    """general_matmul/__init__.py:685-690: int formats are clamped to [-maxq, maxq] and shifted by +maxq."""
    import bitblas_b200 as bitblas
    op = bitblas.Matmul(bitblas.MatmulConfig(M=1, N=16, K=64, A_dtype="float16", W_dtype="int4"), enable_tuning=False)
    w = torch.randint(-20, 20, (16, 64), dtype=torch.int8)
    got = op.transform_weight(w)
    exp = O.transform_weight(w, "int4", "float16", fast_decoding=True)
    assert torch.equal(got, exp)
    # quirk kept from the reference (:711): extra operands are accepted and only the weight is returned
    op2 = bitblas.Matmul(bitblas.MatmulConfig(M=1, N=16, K=64, A_dtype="float16", W_dtype="uint4", with_scaling=True), enable_tuning=False)
    out = op2.transform_weight(torch.randint(0, 16, (16, 64), dtype=torch.int8), scale=torch.ones(16, 1).half())
    assert isinstance(out, torch.Tensor) and tuple(out.shape) == (16, 32)
