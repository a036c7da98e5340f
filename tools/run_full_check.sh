#!/bin/bash
# Final GPU session of a round: the whole -m gpu suite, the bench (N = 1) + reference arm, the ncu launch list of the bench command,
# full-set captures of the two dominant kernels, and a compute-sanitizer pass over a subset of the parity tests.
mkdir -p gpurun_out
if [ "$1" != "nopytest" ]; then
timeout 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -30 > gpurun_out/r2_pytest_gpu.txt
tail -3 gpurun_out/r2_pytest_gpu.txt
fi
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 1500 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
tail -c 400 gpurun_out/r2_bench_reference_arm.json
if [ "$1" != "quick" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --skip-cpu --only x > gpurun_out/r2_launch_run.log 2>&1
tail -2 gpurun_out/r2_launch_run.log
timeout 600 ncu --set full --clock-control none -k regex:gemm_ts -s 4 -c 1 -o gpurun_out/r2_gemm_prof \
  python bench.py --steps 2 --warmup 3 --skip-cpu --only gemm > gpurun_out/r2_gemm_ncu.log 2>&1
tail -2 gpurun_out/r2_gemm_ncu.log
# (gpurun copies back at most 64 MiB: keep the raw-page CSV of the GEMM capture, drop the report)
ncu -i gpurun_out/r2_gemm_prof.ncu-rep --page raw --csv > gpurun_out/r2_gemm_prof_raw.csv 2>/dev/null; rm -f gpurun_out/r2_gemm_prof.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_slab -s 12 -c 2 -o gpurun_out/r2_slab_prof \
  tools/gemv_bench --iters 5 --nocheck 12288x12288 > gpurun_out/r2_slab_ncu.log 2>&1
tail -2 gpurun_out/r2_slab_ncu.log
ncu -i gpurun_out/r2_slab_prof.ncu-rep --page raw --csv > gpurun_out/r2_slab_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_slab_prof.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_slab_prof_source.csv 2>/dev/null
rm -f gpurun_out/r2_slab_prof.ncu-rep
{ echo "# compute-sanitizer memcheck"; timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scatter.py tests/test_gpu_tile.py tests/test_gpu_gemv_slab.py -q -m gpu -x -k "scatter_matches or device_retile or (tiled_matches and 256) or (gemv_slab_parity and 256)" 2>&1 | tail -8;
  echo "# compute-sanitizer racecheck"; timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scatter.py -q -m gpu -x -k "scatter_matches" 2>&1 | grep -v "^=========     Saved\|^=========         Host Frame" | head -120; } > gpurun_out/r2_sanitizer.txt
tail -5 gpurun_out/r2_sanitizer.txt; ls -la gpurun_out; du -sh gpurun_out
fi
