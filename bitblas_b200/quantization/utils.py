"""The weight storage format kept from the reference (bitblas/quantization/utils.py:8-110): little-endian
sub-byte packing (``general_compress``) and the per-int32 LOP3 interleave (``interleave_weight``).

Vectorised numpy on unsigned 32-bit words (the reference's signed-int32 masks overflow under NumPy 2 and its
1-bit/float16 branch drops its result, SURVEY.md §8c defect (i)); bit-exact with the reference's numpy and C++
host functions -- see tests/test_quantization_format.py.  Bulk conversion for real models goes through the C++
/ CUDA routines (bb_compress_host, bb_interleave_host, bb_transform_weight_device).
"""
import numpy as np
import torch
import torch.nn as nn


def gen_quant4(k, n, groupsize=-1):
    """Test helper of the reference (quantization/utils.py:8-51): random fp16 weight -> symmetric 4-bit."""
    maxq = 2**4
    w = torch.randn((k, n), dtype=torch.half, device="cpu")
    original_w = w.clone()
    if groupsize == -1:
        groupsize = k
    w = w.reshape((-1, groupsize, n)).permute(1, 0, 2).reshape((groupsize, -1))
    s = torch.max(torch.abs(w), 0, keepdim=True)[0]
    s *= 2 / maxq
    w = torch.round(w / s).int()
    w += maxq // 2
    w = torch.clamp(w, 0, maxq)
    ref = (w - maxq // 2).half() * s

    def _reshape(t):
        return t.reshape((groupsize, -1, n)).permute(1, 0, 2).reshape((k, n)).contiguous()

    ref = _reshape(ref)
    w = _reshape(w)
    s = s.reshape((-1, n)).contiguous()
    linear = nn.Linear(k, n, bias=False)
    linear.weight.data = ref.t()
    return original_w, linear, s, (w - maxq // 2)


def general_compress(lowprecision_weight, source_bits=4, storage_dtype=np.int8):
    w = np.asarray(lowprecision_weight)
    if w.dtype == np.float16:
        w = w.astype(np.int8)
    epb = 8 // source_bits
    if w.shape[-1] % epb:
        raise ValueError(f"last dimension {w.shape[-1]} is not a multiple of {epb}")
    u = w.astype(np.uint8).astype(np.uint32).reshape(*w.shape[:-1], w.shape[-1] // epb, epb)
    shifts = np.arange(epb, dtype=np.uint32) * np.uint32(source_bits)
    packed = np.bitwise_or.reduce((u << shifts) & np.uint32(0xFF), axis=-1).astype(np.uint8)
    return packed.view(np.int8).view(storage_dtype)


_NIBBLE_MAP_1B_F16 = (0, 2, 4, 6, 1, 3, 5, 7)
_NIBBLE_MAP_1B_I8 = (0, 4, 2, 6, 1, 5, 3, 7)


def _field_bitpos(o, nbits, bits_stride):
    num_groups = 32 // bits_stride
    pos = (o % num_groups) * bits_stride + (o // num_groups) * nbits
    if nbits == 2 and bits_stride == 16:
        byte = pos >> 3
        pos += 8 if byte == 1 else (-8 if byte == 2 else 0)
    elif nbits == 1 and bits_stride == 16:
        pos = _NIBBLE_MAP_1B_F16[pos >> 2] * 4 + (pos & 3)
    elif nbits == 1 and bits_stride == 8:
        pos = _NIBBLE_MAP_1B_I8[pos >> 2] * 4 + (pos & 3)
    return pos


def interleave_weight(qweight, nbits=4, target_dtype="float16"):
    assert target_dtype in ["float16", "bfloat16", "int8"]
    q = np.ascontiguousarray(qweight).view(np.uint32)
    new = np.zeros_like(q)
    bits_stride = 8 if target_dtype == "int8" else 16
    mask = np.uint32((1 << nbits) - 1)
    for o in range(32 // nbits):
        new |= ((q >> np.uint32(nbits * o)) & mask) << np.uint32(_field_bitpos(o, nbits, bits_stride))
    return new.view(np.int8)
