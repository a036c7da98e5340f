"""bench.py's bookkeeping: the algorithmic-byte formula behind every GB/s figure, and that both arms describe the
same workload."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_formula():
    b = _bench()
    N = K = 12288
    # W + scale + packed zeros + A + C for uint4, g = 128, quantized zeros (the figure DESIGN.md and the verdict quote)
    assert b.gemv_bytes(N, K) == N * K // 2 + N * (K // 128) * 2 + (K // 128) * N // 2 + K * 2 + N * 2 == 78495744
    assert b.gemv_bytes(N, K, zeros="none") == N * K // 2 + N * (K // 128) * 2 + K * 2 + N * 2
    assert b.gemv_bytes(N, K, zeros="original") == N * K // 2 + 2 * N * (K // 128) * 2 + K * 2 + N * 2
    step = sum(b.gemv_bytes(n, k) for n, k in b.GEMV_SHAPES)
    assert step == 357597184          # the 4-projection step both arms and every world size are normalised by


def test_both_arms_share_the_config():
    b = _bench()
    assert b.CONFIG["shapes_NK"] == b.GEMV_SHAPES and b.CONFIG["M"] == 1 and b.CONFIG["group_size"] == b.GROUP
    assert b.CONFIG["W_dtype"] == "uint4" and b.CONFIG["zeros_mode"] == "quantized"
