#!/bin/bash
# table / fp8 weight formats on the fast kernels: parity + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "table or generic or fp8 or e4m3" 2>&1 | tail -15 > gpurun_out/r2_formats_pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu --only formats 2> gpurun_out/r2_formats.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d.get('table_formats'))); print(json.dumps(d.get('fp8_weights')))" > gpurun_out/r2_formats_bench.txt 2>&1
cat gpurun_out/r2_formats_pytest.txt gpurun_out/r2_formats_bench.txt; tail -3 gpurun_out/r2_formats.err
