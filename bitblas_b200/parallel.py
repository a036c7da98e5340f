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
                 pipeline_chunks: Optional[int] = None, fused_gather: Optional[bool] = None, fused_ring: int = 2,
                 clone_output: bool = False, **linear_kwargs):
        super().__init__()
        # fused_gather: the matmul kernel's epilogue stores this rank's column slice straight into every rank's output
        # (peer-mapped symmetric memory over NVLink/NVSwitch) -- no separate collective, the transfer overlaps the math
        # tile by tile.  None = decided COLLECTIVELY at the first CUDA forward (every rank must be able to use it, else all
        # ranks take the NCCL all-gather; a rank-local choice would leave peers waiting in a rendezvous / barrier).
        # LIFETIME of the fused result: forward() returns a view into a ring of `fused_ring` symmetric buffers; it stays
        # valid until `fused_ring - 1` further forward() calls of this layer have been issued (the call after that lets the
        # peers overwrite it).  Pass clone_output=True (one extra copy) or a deeper ring to hold results longer.
        if fused_ring < 2:
            raise ValueError("fused_ring must be >= 2 (a rank may only overwrite a buffer after the next call's barrier)")
        self.fused_gather = fused_gather
        self.fused_ring = int(fused_ring)
        self.clone_output = bool(clone_output)
        self._fused_decided = fused_gather is False
        self._symm = {}       # (rows capacity, dtype) -> [ring of (tensor, handle, peer pointer array), next index]
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
            for _ in range(self.fused_ring):  # ring: one barrier per call is enough (see _forward_fused)
                t = symm_mem.empty((cap, self.out_features), dtype=dtype, device=device)
                h = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in h.buffer_ptrs]
                arr = (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(p) for p in ptrs])   # built once per buffer
                pairs.append((t, h, arr))
            self._symm[key] = [pairs, 0]
        entry = self._symm[key]
        t, h, arr = entry[0][entry[1]]
        entry[1] = (entry[1] + 1) % self.fused_ring
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
        return buf[:m].clone() if self.clone_output else buf[:m]

    def _fused_supported_here(self, x2: torch.Tensor) -> bool:
        """can THIS rank run the fused path for this call: symmetric memory importable and a kernel family with the scatter
        epilogue (everything but the generic SIMT kernel) dispatched for this m."""
        if not x2.is_cuda or self.local.consistent:
            return False
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
        except Exception:
            return False
        return self.local.bitblas_matmul.kernel_for(int(x2.shape[0])) != "generic_simt"

    def _decide_fused(self, x2: torch.Tensor) -> None:
        """one collective decision per layer (MIN over ranks of the local capability), taken at the first CUDA forward."""
        ok = torch.tensor([1 if self._fused_supported_here(x2) else 0], dtype=torch.int32, device=x2.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        agreed = bool(int(ok.item()))
        if self.fused_gather is True and not agreed:
            raise RuntimeError("fused_gather=True was requested but at least one rank cannot use it")
        self.fused_gather = agreed
        self._fused_decided = True

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.world == 1 or not self.gather_output:
            return self.local(x)
        if x.is_cuda and not self._fused_decided:
            self._decide_fused(x.reshape(-1, x.shape[-1]))
        if x.is_cuda and self.fused_gather:
            x2 = x.reshape(-1, x.shape[-1])
            # per-call probe (the dispatcher's choice depends on m): an m that lands on the generic kernel takes the NCCL path
            # for this call only -- identically on every rank, since all ranks see the same m and the same shard shape
            if self.local.bitblas_matmul.kernel_for(int(x2.shape[0])) != "generic_simt":
                return self._forward_fused(x2).reshape(*x.shape[:-1], self.out_features)
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
