#!/bin/bash
# round-2 hardware pass on ONE GPU: GPU tests, the bench line, the ncu launch list of the bench command,
# phase timers of the C3 pipeline.  Everything lands under gpurun_out/r02_* (summaries are copied to profiles/).
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  echo "== pytest -m gpu"
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02_gpu_tests.txt
fi
echo "== bench N=1"
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_FLAGS:-} > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 6000 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
echo "== pipeline phase timers"
timeout 400 python tools/run_pipeline.py C3 2 --patches 2>&1 | grep -v "^trace" | tail -40 | tee gpurun_out/r02_pipeline_c3.txt
if [ "${SKIP_NCU:-0}" != 1 ]; then
  echo "== ncu launch list (bench command, 2 steps)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv \
      python bench.py --steps 2 --warmup 1 --no-verify --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
  tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-400
  wc -l gpurun_out/r02_launches_bench.csv
fi
