#!/bin/bash
# GPU session: the whole -m gpu suite, then the bench (N = 1) with its launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=12 2>&1 | tail -40 > gpurun_out/r2_pytest_gpu.txt
tail -5 gpurun_out/r2_pytest_gpu.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 3000 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
tail -c 600 gpurun_out/r2_bench_reference_arm.json
