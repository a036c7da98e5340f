"""Column-parallel (N-sharded) low-bit linear over the GPUs of one node.

The reference has no distributed code at all (SURVEY.md §2a "Collectives: none"); this is the optional multi-GPU
path BASELINE.json's north_star asks for.  The matmul shards naturally along N: output column n depends only on
row n of W / scales / zeros (column n of quantized zeros / bias), the per-int32 pack + interleave runs along K so a
row shard never splits a storage word.  Rank r of G owns rows [r*N/G, (r+1)*N/G); activations are replicated (as
in tensor-parallel inference); one exchange step -- an all-gather of the [m, N/G] partial outputs -- rebuilds C.

One process per GPU, ``torch.distributed`` (NCCL over NVLink5/NVSwitch; gloo on CPU for the host-logic tests).
For large m the all-gather is pipelined against the matmul in row chunks on a side stream so the transfer of
chunk i overlaps the tcgen05 kernel of chunk i+1.
"""
from __future__ import annotations

from typing import Optional

import ctypes

import torch
import torch.distributed as dist
import torch.nn as nn

from .module import Linear


def shard_bounds(n: int, rank: int, world: int, multiple: int = 16):
    if n % world:
        raise ValueError(f"N={n} is not divisible by world size {world}")
    per = n // world
    if per % multiple:
        raise ValueError(f"N/world = {per} must be a multiple of {multiple}")
    return rank * per, (rank + 1) * per


def shard_quantized_params(qweight, scales, zeros, bias, *, rank: int, world: int, bits: int, zeros_mode: str = "original"):
    """Slice full-size BitBLAS-layout parameters for one rank (works on CPU or CUDA tensors)."""
    N = qweight.shape[0]
    lo, hi = shard_bounds(N, rank, world)
    out = {"qweight": qweight[lo:hi].contiguous()}
    if scales is not None:
        out["scales"] = scales[lo:hi].contiguous()
    if zeros is not None:
        if zeros_mode == "quantized":
            epb = 8 // bits
            if lo % epb or hi % epb:
                raise ValueError("shard boundary splits a packed zero-point byte")
            out["zeros"] = zeros[:, lo // epb:hi // epb].contiguous()
        else:
            out["zeros"] = zeros[lo:hi].contiguous()
    if bias is not None:
        out["bias"] = bias[lo:hi].contiguous()
    return out


class ColumnParallelLinear(nn.Module):
    """y = gather_N( x @ dequant(W_r)^T (+ b_r) ) with W sharded by rows (output features) across the group."""

    def __init__(self, in_features: int, out_features: int, *, process_group=None, gather_output: bool = True,
                 pipeline_chunks: Optional[int] = None, fused_gather: Optional[bool] = None, **linear_kwargs):
        super().__init__()
        # fused_gather: the matmul kernel's epilogue stores this rank's column slice straight into every rank's output
        # (peer-mapped symmetric memory over NVLink/NVSwitch) -- no separate collective, the transfer overlaps the math
        # tile by tile.  None = use it when torch symmetric memory is available on CUDA, else NCCL all-gather.
        self.fused_gather = fused_gather
        self._symm = {}       # (rows capacity, dtype) -> [two (tensor, handle) pairs, next index]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.in_features = in_features
        self.out_features = out_features
        lo, hi = shard_bounds(out_features, self.rank, self.world)
        self.n_lo, self.n_hi = lo, hi
        self.gather_output = gather_output
        self.pipeline_chunks = pipeline_chunks
        self.local = Linear(in_features, hi - lo, **linear_kwargs)
        self._comm_stream = None

    def load_full_params(self, qweight, scales=None, zeros=None, bias=None):
        """Take full-size (already transformed) parameters and keep this rank's shard."""
        op = self.local.bitblas_matmul
        sh = shard_quantized_params(qweight, scales, zeros, bias, rank=self.rank, world=self.world, bits=op.bit,
                                    zeros_mode=op.config.zeros_mode)
        dev = self.local.qweight.device
        self.local.qweight = sh["qweight"].to(dev)
        if "scales" in sh:
            self.local.scales = sh["scales"].to(dev)
        if "zeros" in sh:
            self.local.zeros = sh["zeros"].to(dev)
        if "bias" in sh:
            self.local.bias = sh["bias"].to(dev)
        self.local.q_params = None

    def _chunks_for(self, m: int) -> int:
        if self.pipeline_chunks is not None:
            return max(1, min(self.pipeline_chunks, m))
        return max(1, min(8, m // 512))

    def _symm_buffers(self, m: int, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        cap = 1
        while cap < m:
            cap *= 2
        key = (cap, dtype)
        if key not in self._symm:
            pairs = []
            for _ in range(2):  # double buffered: one barrier per call is enough (see forward)
                t = symm_mem.empty((cap, self.out_features), dtype=dtype, device=device)
                h = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in h.buffer_ptrs]
                arr = (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(p) for p in ptrs])   # built once per buffer
                pairs.append((t, h, arr))
            self._symm[key] = [pairs, 0]
        entry = self._symm[key]
        t, h, arr = entry[0][entry[1]]
        entry[1] ^= 1
        return t, h, arr

    def _forward_fused(self, x2: torch.Tensor) -> torch.Tensor:
        m = x2.shape[0]
        op = self.local.bitblas_matmul
        out_dtype = getattr(torch, op.out_dtype)
        buf, hdl, peer_arr = self._symm_buffers(m, out_dtype, x2.device)
        lin = self.local
        op.forward_scatter(x2.contiguous(), lin.qweight, scale=lin.scales if op.with_scaling else None,
                           zeros=lin.zeros if op.with_zeros else None, bias=lin.bias if op.with_bias else None,
                           peer_ptrs=peer_arr, ldc=self.out_features, col_offset=self.n_lo)
        # every rank's slice has landed in every buffer once all ranks passed this point.  Buffers alternate between calls:
        # a rank can only overwrite buffer b again after the NEXT call's barrier, which every peer reaches after its
        # (stream-ordered) reads of this call's result.
        hdl.barrier(channel=0)
        return buf[:m]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.world == 1 or not self.gather_output:
            return self.local(x)
        if x.is_cuda and self.fused_gather is not False and not self.local.consistent:
            try:
                out = self._forward_fused(x.reshape(-1, x.shape[-1]))
                self.fused_gather = True
                return out.reshape(*x.shape[:-1], self.out_features)
            except Exception as e:  # symmetric memory unavailable / kernel family without a scatter epilogue
                if self.fused_gather is True:
                    raise
                import logging
                logging.getLogger(__name__).warning("fused column-parallel gather unavailable (%s); using NCCL all-gather", e)
                self.fused_gather = False
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        m = x2.shape[0]
        per = self.n_hi - self.n_lo
        out_dtype = getattr(torch, self.local.bitblas_matmul.out_dtype)
        full = torch.empty((m, self.out_features), dtype=out_dtype, device=x.device)
        nchunks = self._chunks_for(m)
        if x.is_cuda and self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=x.device)
        bounds = [(i * m) // nchunks for i in range(nchunks + 1)]
        staged = []
        for i in range(nchunks):
            r0, r1 = bounds[i], bounds[i + 1]
            if r1 == r0:
                continue
            part = self.local(x2[r0:r1]).contiguous()         # [rows, N/G] on the compute stream
            gathered = torch.empty((self.world * (r1 - r0), per), dtype=out_dtype, device=x.device)
            if x.is_cuda:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(x.device))
                with torch.cuda.stream(self._comm_stream):
                    self._comm_stream.wait_event(ev)
                    dist.all_gather_into_tensor(gathered, part, group=self.group)
                    part.record_stream(self._comm_stream)
            else:
                dist.all_gather_into_tensor(gathered, part, group=self.group)
            staged.append((r0, r1, gathered))
        if x.is_cuda:
            torch.cuda.current_stream(x.device).wait_stream(self._comm_stream)
        for r0, r1, gathered in staged:
            # [G, rows, N/G] -> [rows, G, N/G] == rows of the full output
            full[r0:r1].view(r1 - r0, self.world, per).copy_(gathered.view(self.world, r1 - r0, per).permute(1, 0, 2))
        return full.reshape(*lead, self.out_features)


__all__ = ["ColumnParallelLinear", "shard_quantized_params", "shard_bounds"]
