#!/bin/bash
# N-GPU session (gpurun --gpus N): multi-GPU parity test (optional), then the bench at N with the fused path
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/r2_multi_${N}gpu.txt
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT
if [ "$2" != "benchonly" ]; then
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -s 2>&1 | grep -v Warning | tail -8 >> $OUT
fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $N --steps 30 --warmup 5 --only gemm > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
python - <<PY >> $OUT
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=$N value", d["value"], "ms_per_step", d["ms_per_step"], d["config_detail"]["parallelism"], d["step_method"])
    print("   gemm", d["gemm"]["ms"], d["gemm"]["TFLOPS"], "small", [(r["M"], r["us"]) for r in d["gemm_small_m"]])
    print("   shapes", [(r["N"], r["K"], r["us"]) for r in d["gemv_shapes"]])
except Exception as e:
    print("N=$N FAILED", e)
    print(open("gpurun_out/r2_bench_${N}gpu.err").read()[-2500:])
PY
cat $OUT
