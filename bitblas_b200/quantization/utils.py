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
    """Random fp16 weight [k, n] quantised symmetrically to 4 bits per group of `groupsize` input features -- the helper the
    reference's tests and integrations import from here (quantization/utils.py:8-51).  Returns, like the reference,
    (original weight, nn.Linear holding the dequantised weight, scales [k/g, n], signed integer weight [k, n]); the integer
    range keeps the reference's clamp to [0, 16] before the -8 shift."""
    levels = 16
    g = k if groupsize == -1 else groupsize
    original_w = torch.randn((k, n), dtype=torch.half, device="cpu")
    grouped = original_w.view(k // g, g, n)
    scales = grouped.abs().amax(dim=1, keepdim=True) * (2 / levels)              # [k/g, 1, n]
    signed = torch.clamp(torch.round(grouped / scales).int() + levels // 2, 0, levels) - levels // 2
    linear = nn.Linear(k, n, bias=False)
    linear.weight.data = (signed.half() * scales).view(k, n).t()
    return original_w, linear, scales.view(k // g, n).contiguous(), signed.view(k, n)


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
