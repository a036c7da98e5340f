#!/bin/bash
# single-GPU check of the scatter epilogue + A/B of the staged GEMM epilogue on the plain path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scatter.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_scatter_pytest.txt
for st in 0 1; do
  BB_TS_STAGED=$st timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --only gemm 2> gpurun_out/r2_staged${st}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('staged=$st', 'gemm', d['gemm']['ms'], d['gemm']['TFLOPS'], [ (r['M'], r['us']) for r in d['gemm_small_m']], 'step', d['ms_per_step'], d['value'], d['step_method'])
" > gpurun_out/r2_staged${st}.txt 2>&1
done
cat gpurun_out/r2_scatter_pytest.txt gpurun_out/r2_staged0.txt gpurun_out/r2_staged1.txt
