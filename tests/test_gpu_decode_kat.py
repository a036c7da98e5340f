"""GPU known-answer tests for the in-register LOP3 decode (bb_debug_decode), the analogue of the reference's
gtest DecodeTest.* suite (testing/cpp/lop3_type_conversion/lowprecision_to_float16.cu:51-101, lowprecision_to_int8.cu):
values -> compress -> interleave -> device decode -> exact equality.  When oracle/_ref is present the same packed
words are also decoded by the REFERENCE's device functions (oracle/ref_shim.cu) and compared bit for bit."""
import ctypes
import os

import numpy as np
import pytest
import torch

import bitblas_oracle as O
from bitblas_b200 import _lib

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libbitblas_ref.so")


def _pack(values, bits, tgt, layout_il=True):
    packed = O.general_compress(values, bits)
    if layout_il:
        packed = O.interleave_weight(packed, bits, tgt)
    return torch.from_numpy(np.ascontiguousarray(packed)).cuda()


@pytest.mark.parametrize("bits", [4, 2])
@pytest.mark.parametrize("signed", [0, 1])
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("il", [True, False])
def test_decode_to_16bit_exact(bits, signed, kind, il):
    lib = _lib.load()
    _lib.ensure_init(0)
    rng = np.random.RandomState(0)   # srand(0) in the reference KATs
    vals = rng.randint(0, 2**bits, size=(1, 4096)).astype(np.int8)
    dev = _pack(vals, bits, "float16", il)
    out = torch.empty(4096, dtype=torch.float16 if kind == 0 else torch.bfloat16, device="cuda")
    layout = _lib.BB_LAYOUT_INTERLEAVED_16 if il else _lib.BB_LAYOUT_COMPRESSED
    _lib.check(lib.bb_debug_decode(kind, bits, signed, layout, dev.data_ptr(), out.data_ptr(), dev.numel() // 4, 0))
    torch.cuda.synchronize()
    expect = vals.reshape(-1).astype(np.int32) - (2 ** (bits - 1) if signed else 0)
    assert np.array_equal(out.float().cpu().numpy().astype(np.int32), expect)


@pytest.mark.parametrize("bits", [4, 2])
@pytest.mark.parametrize("signed", [0, 1])
def test_decode_to_int8_exact(bits, signed):
    lib = _lib.load()
    _lib.ensure_init(0)
    rng = np.random.RandomState(0)
    vals = rng.randint(0, 2**bits, size=(1, 4096)).astype(np.int8)
    dev = _pack(vals, bits, "int8")
    out = torch.empty(4096, dtype=torch.int8, device="cuda")
    _lib.check(lib.bb_debug_decode(2, bits, signed, _lib.BB_LAYOUT_INTERLEAVED_8, dev.data_ptr(), out.data_ptr(), dev.numel() // 4, 0))
    torch.cuda.synchronize()
    expect = vals.reshape(-1).astype(np.int32) - (2 ** (bits - 1) if signed else 0)
    assert np.array_equal(out.cpu().numpy().astype(np.int32), expect)


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")
@pytest.mark.parametrize("bits,kind_ref", [(4, 0), (2, 2)])
def test_decode_matches_reference_device_functions(bits, kind_ref):
    """unsigned decode vs decode_i4u_to_f16 / decode_i2u_to_f16 of fast_decoding.hpp run on this GPU."""
    lib = _lib.load()
    _lib.ensure_init(0)
    ref = ctypes.CDLL(REF_SO)
    ref.ref_decode_f16.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 4
    rng = np.random.RandomState(5)
    vals = rng.randint(0, 2**bits, size=(1, 8192)).astype(np.int8)
    dev = _pack(vals, bits, "float16")
    mine = torch.empty(8192, dtype=torch.float16, device="cuda")
    theirs = torch.empty(8192, dtype=torch.float16, device="cuda")
    _lib.check(lib.bb_debug_decode(0, bits, 0, _lib.BB_LAYOUT_INTERLEAVED_16, dev.data_ptr(), mine.data_ptr(), dev.numel() // 4, 0))
    assert ref.ref_decode_f16(kind_ref, dev.data_ptr(), theirs.data_ptr(), 8192 // 8, None, None, None, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(mine, theirs)
    assert np.array_equal(theirs.cpu().numpy().astype(np.int32), vals.reshape(-1))


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built")
def test_int8_decode_matches_reference_device_functions():
    lib = _lib.load()
    _lib.ensure_init(0)
    ref = ctypes.CDLL(REF_SO)
    ref.ref_decode_i8.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    rng = np.random.RandomState(6)
    vals = rng.randint(0, 4, size=(1, 4096)).astype(np.int8)
    dev = _pack(vals, 2, "int8")
    mine = torch.empty(4096, dtype=torch.int8, device="cuda")
    theirs = torch.empty(4096, dtype=torch.int8, device="cuda")
    _lib.check(lib.bb_debug_decode(2, 2, 0, _lib.BB_LAYOUT_INTERLEAVED_8, dev.data_ptr(), mine.data_ptr(), dev.numel() // 4, 0))
    assert ref.ref_decode_i8(2, dev.data_ptr(), theirs.data_ptr(), 4096 // 16, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(mine, theirs)


@pytest.mark.parametrize("bits", [4, 2, 1])
@pytest.mark.parametrize("tgt", [0, 16, 8])
def test_device_weight_transform_matches_host(bits, tgt):
    from bitblas_b200.ops.general_matmul import WeightTransform
    rng = np.random.RandomState(7)
    w = torch.from_numpy(rng.randint(0, 2**bits, size=(64, 512)).astype(np.int8))
    wt = WeightTransform(bits, tgt)
    assert torch.equal(wt(w.cuda()).cpu(), wt(w))
