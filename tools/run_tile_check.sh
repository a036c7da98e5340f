#!/bin/bash
# slab-tiled weight storage (propagate_b) + helper-warp backoff A/B
mkdir -p gpurun_out
OUT=gpurun_out/r2_tile_bench.txt
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_gemv_slab.py -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_tile_pytest.txt
echo "== row-major" > $OUT
timeout 300 tools/gemv_bench --iters 200 --cfg 4,2,1 --cfg 4,2,2 12288x12288 8192x8192 28672x8192 8192x28672 >> $OUT 2>&1
echo "== slab-tiled W (propagate_b)" >> $OUT
timeout 300 tools/gemv_bench --iters 200 --tile --cfg 4,2,1 12288x12288 8192x8192 >> $OUT 2>&1
echo "== feed only (BB_GS_DBG=21 / 1)" >> $OUT
BB_GS_DBG=21 timeout 120 tools/gemv_bench --iters 200 --nocheck --cfg 4,2,1 12288x12288 >> $OUT 2>&1
BB_GS_DBG=1 timeout 120 tools/gemv_bench --iters 200 --nocheck --cfg 4,2,1 12288x12288 >> $OUT 2>&1
tail -4 gpurun_out/r2_tile_pytest.txt; cat $OUT
