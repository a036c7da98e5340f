import os, sys, torch
sys.path.insert(0, "/root/repo")
import bitblas_b200 as bb
dev="cuda"; N=K=12288
cfg = bb.MatmulConfig(M=1, N=N, K=K, A_dtype="int8", W_dtype="int4", accum_dtype="int32", out_dtype="int32")
op = bb.Matmul(cfg, enable_tuning=False)
Ws = [torch.randint(-128, 128, (N, K // 2), dtype=torch.int8, device=dev) for _ in range(6)]
A = torch.randint(-128, 128, (1, K), dtype=torch.int8, device=dev)
out = torch.empty(1, N, dtype=torch.int32, device=dev)
for i in range(8): op.forward(A, Ws[i % 6], output=out)
torch.cuda.synchronize()
