#!/bin/bash
# GPU session: slab GEMV correctness + timing + ncu (run under gpurun from the repo root)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r2_slab_bench.txt
echo "== small shape sanity" >> gpurun_out/r2_slab_bench.txt
timeout 120 tools/gemv_bench --iters 20 1024x2048 4112x6144 >> gpurun_out/r2_slab_bench.txt 2>&1 || echo "SANITY FAILED rc=$?" >> gpurun_out/r2_slab_bench.txt
echo "== sweep" >> gpurun_out/r2_slab_bench.txt
timeout 300 tools/gemv_bench --iters 200 --cfg 4,2 --cfg 3,2 --cfg 3,3 --cfg 6,1 --cfg 2,2 12288x12288 8192x8192 28672x8192 8192x28672 >> gpurun_out/r2_slab_bench.txt 2>&1 || echo "SWEEP FAILED rc=$?" >> gpurun_out/r2_slab_bench.txt
echo "== no PDL" >> gpurun_out/r2_slab_bench.txt
BB_PDL=0 timeout 120 tools/gemv_bench --iters 200 --nocheck --cfg 4,2,-1 12288x12288 8192x8192 >> gpurun_out/r2_slab_bench.txt 2>&1
echo "== old kernel (gemv_mma)" >> gpurun_out/r2_slab_bench.txt
timeout 120 tools/gemv_bench --iters 200 --nocheck --kernel 2 12288x12288 8192x8192 >> gpurun_out/r2_slab_bench.txt 2>&1
tail -40 gpurun_out/r2_slab_bench.txt
timeout 900 python -m pytest tests/test_gpu_gemv_slab.py -x -q 2>&1 | tail -15 > gpurun_out/r2_slab_pytest.txt
cat gpurun_out/r2_slab_pytest.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_slab -s 12 -c 2 -o gpurun_out/r2_slab_prof tools/gemv_bench --iters 5 --nocheck 12288x12288 > gpurun_out/r2_slab_ncu.log 2>&1
tail -3 gpurun_out/r2_slab_ncu.log
