"""CPU, world_size 2, gloo: the host logic of the column-parallel path (sharding + all-gather re-assembly).

There is no CPU matmul in the product, so each rank's local operator is replaced by a stand-in that evaluates
the ORACLE on that rank's shard (test infrastructure only) -- what is under test is shard_quantized_params,
the chunked all-gather and the [G, rows, N/G] -> [rows, N] re-assembly of ColumnParallelLinear.forward."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, chunks, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bitblas_oracle as O
        import helpers as H
        from bitblas_b200.parallel import ColumnParallelLinear, shard_quantized_params
        M, N, K, g = 12, 64, 256, 128
        for mode in ("original", "quantized"):
            case = H.make_case(M, N, K, W_dtype="uint4", group_size=g, with_scaling=True, with_zeros=True, zeros_mode=mode,
                               with_bias=True, seed=7)
            full_ref = H.oracle_output(case)
            layer = ColumnParallelLinear(K, N, bias=True, A_dtype="float16", W_dtype="uint4", group_size=g, with_scaling=True,
                                         with_zeros=True, zeros_mode=mode, enable_tuning=False, pipeline_chunks=chunks)
            stored = layer.local.bitblas_matmul.weight_transform(case["fields"].to(torch.int8))
            layer.load_full_params(stored, case["scale"], case["zeros"], case["bias"])
            lo, hi = layer.n_lo, layer.n_hi
            assert (lo, hi) == (rank * N // world, (rank + 1) * N // world)
            assert torch.equal(layer.local.qweight, stored[lo:hi])
            sh = shard_quantized_params(stored, case["scale"], case["zeros"], case["bias"], rank=rank, world=world, bits=4,
                                        zeros_mode=mode)
            shard_fields = torch.from_numpy(O.general_decompress(
                O.deinterleave_weight(sh["qweight"].numpy(), 4, "float16"), 4)).to(torch.int32)
            assert torch.equal(shard_fields, case["fields"][lo:hi])

            def fake_local(x, _sh=sh, _f=shard_fields, _mode=mode):
                return O.matmul_dequant(x, _f, W_dtype="uint4", group_size=g, with_scaling=True, with_zeros=True,
                                        zeros_mode=_mode, scale=_sh["scales"], zeros=_sh["zeros"], bias=_sh["bias"])

            layer.local.forward = fake_local
            out = layer(case["A"])
            assert out.shape == (M, N)
            assert torch.equal(out, full_ref), f"rank {rank} mode {mode}"
            out3 = layer(case["A"].reshape(3, 4, K))
            assert torch.equal(out3.reshape(M, N), full_ref)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chunks", [1, 3])
def test_column_parallel_world2(chunks):
    world = 2
    port = 29500 + os.getpid() % 1000 + chunks
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, chunks, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_shard_bounds_errors():
    from bitblas_b200.parallel import shard_bounds
    assert shard_bounds(12288, 3, 8) == (4608, 6144)
    with pytest.raises(ValueError):
        shard_bounds(100, 0, 8)
    with pytest.raises(ValueError):
        shard_bounds(8 * 24, 0, 8)
