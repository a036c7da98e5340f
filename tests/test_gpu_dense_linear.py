"""The dense ("consistent", A_dtype == W_dtype) side of bitblas.Matmul / bitblas.Linear: a library GEMM behind the same API
(reference: bitblas/ops/general_matmul/__init__.py:33-51,568-580; bitblas/module/__init__.py:120,298-304; tests:
testing/python/module/test_bitblas_linear.py:23-60 compare against torch.nn.Linear)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,in_features,out_features,bias", [(1, 1024, 1024, False), (1, 1024, 1024, True),
                                                              (1024, 1024, 1024, True), (7, 512, 768, True)])
def test_linear_float16_matches_nn_linear(m, in_features, out_features, bias):
    import bitblas_b200 as bitblas
    torch.manual_seed(0)
    ref = torch.nn.Linear(in_features, out_features, bias=bias).half().cuda()
    lin = bitblas.Linear(in_features, out_features, bias=bias, A_dtype="float16", W_dtype="float16",
                         accum_dtype="float16", out_dtype="float16", opt_M=m, enable_tuning=False)
    assert lin.consistent and lin.bitblas_matmul.lib is None
    lin.load_and_transform_weight(ref.weight.data.clone())
    if bias:
        lin.bias = ref.bias.data.clone()
    x = (torch.rand(m, in_features) - 0.5).half().cuda()
    with torch.no_grad():
        torch.testing.assert_close(lin(x), ref(x), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("m,K", [(1, 1024), (16, 4096), (17, 4096), (300, 16384)])
def test_matmul_int8_dense_is_exact_int32(m, K):
    """INT8 x INT8 with int32 accumulation must be exact where an fp32 accumulation is not: all-extreme operands make every
    partial sum exceed 2^24 after ~1040 terms (K = 16384: 2.7e8)."""
    import bitblas_b200 as bitblas
    N = 256
    cfg = bitblas.MatmulConfig(M=m, N=N, K=K, A_dtype="int8", W_dtype="int8", accum_dtype="int32", out_dtype="int32")
    op = bitblas.Matmul(cfg, enable_tuning=False)
    g = torch.Generator().manual_seed(1)
    A = torch.randint(-128, 128, (m, K), generator=g, dtype=torch.int8)
    W = torch.randint(-128, 128, (N, K), generator=g, dtype=torch.int8)
    A[0] = 127 - (torch.arange(K) % 2).to(torch.int8)   # one row of 127 / 126 against a row of -128s and one of +127s
    W[0] = -128
    W[1] = 127
    ref = A.to(torch.int64) @ W.to(torch.int64).t()
    got = op(A.cuda(), W.cuda())
    assert got.dtype == torch.int32 and torch.equal(got.cpu().to(torch.int64), ref)
    if K >= 4096:
        assert int(ref[0, 0].abs()) > 2**24


def test_matmul_bfloat16_dense():
    import bitblas_b200 as bitblas
    cfg = bitblas.MatmulConfig(M=64, N=512, K=1024, A_dtype="bfloat16", W_dtype="bfloat16", accum_dtype="float32",
                               out_dtype="bfloat16", with_bias=True)
    op = bitblas.Matmul(cfg, enable_tuning=False)
    torch.manual_seed(2)
    A = (torch.rand(64, 1024) - 0.5).bfloat16().cuda()
    W = (torch.rand(512, 1024) - 0.5).bfloat16().cuda()
    b = torch.rand(512).bfloat16().cuda()
    ref = (A.float() @ W.float().t()).bfloat16() + b
    torch.testing.assert_close(op(A, W, bias=b).float(), ref.float(), rtol=2e-2, atol=2e-2)
