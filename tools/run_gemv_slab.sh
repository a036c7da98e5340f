#!/bin/bash
# GPU session: slab GEMV correctness + timing + ncu (run under gpurun from the repo root)
mkdir -p gpurun_out
OUT=gpurun_out/r2_slab_bench.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $OUT
echo "== sweep" >> $OUT
timeout 300 tools/gemv_bench --iters 200 --cfg 4,2 --cfg 8,4 --cfg 6,2 12288x12288 8192x8192 >> $OUT 2>&1 || echo "SWEEP FAILED rc=$?" >> $OUT
for dbg in 1 5 21; do
  echo "== BB_GS_DBG=$dbg (1: no arithmetic, 4: no sums, 16: no parameter conversion)" >> $OUT
  BB_GS_DBG=$dbg timeout 120 tools/gemv_bench --iters 200 --nocheck --cfg 4,2 --cfg 8,4 12288x12288 >> $OUT 2>&1
done
echo "== no PDL" >> $OUT
BB_PDL=0 timeout 120 tools/gemv_bench --iters 200 --nocheck --cfg 4,2 12288x12288 >> $OUT 2>&1
cat $OUT
if [ "$1" == "full" ]; then
timeout 600 python -m pytest tests/test_gpu_gemv_slab.py -x -q 2>&1 | tail -15 > gpurun_out/r2_slab_pytest.txt
cat gpurun_out/r2_slab_pytest.txt
fi
if [ "$1" != "quick" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_slab -s 12 -c 2 -o gpurun_out/r2_slab_prof tools/gemv_bench --iters 5 --nocheck 12288x12288 > gpurun_out/r2_slab_ncu.log 2>&1
tail -3 gpurun_out/r2_slab_ncu.log
fi
