"""W4A8 / W2A8 streaming-kernel timing (IMMA path), to calibrate how far an integer-MMA formulation of the m=1 path can go."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bitblas_b200 as bb
dev = "cuda"
for wd, bits in (("int4", 4), ("int2", 2)):
    for (N, K) in [(12288, 12288), (8192, 8192), (28672, 8192), (8192, 28672)]:
        cfg = bb.MatmulConfig(M=1, N=N, K=K, A_dtype="int8", W_dtype=wd, accum_dtype="int32", out_dtype="int32")
        op = bb.Matmul(cfg, enable_tuning=False)
        nb = N * K * bits // 8
        ncopy = max(3, int(400e6 // nb) + 1)
        Ws = [torch.randint(-128, 128, (N, K * bits // 8), dtype=torch.int8, device=dev) for _ in range(ncopy)]
        A = torch.randint(-128, 128, (1, K), dtype=torch.int8, device=dev)
        out = torch.empty(1, N, dtype=torch.int32, device=dev)
        for w in Ws: op.forward(A, w, output=out)
        torch.cuda.synchronize()
        reps = 60
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(reps): op.forward(A, Ws[i % ncopy], output=out)
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) / reps * 1000
        print(f"{wd} N={N} K={K} m=1 {op.kernel_for(1)} {us:.1f} us  {nb / us / 1e3:.0f} GB/s", flush=True)
        del Ws
