#!/bin/bash
# round-2 GPU pass B (1 GPU): k_tree with cp.async staging, forest phase timers, multi-rank paths on one GPU, bench with verify
set -u
mkdir -p gpurun_out
echo "== gpu tests (MRF parity + multi-rank on one GPU)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi_rank.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/b_tests.txt
echo "== C3 pipeline, forest phase timers"
B2TEX_FOREST_TIMING=1 timeout 400 python tools/run_pipeline.py C3 2 2>&1 | grep -v "^trace" | tail -45 | tee gpurun_out/b_pipeline_c3.txt
echo "== smem / group sweep"
for kb in 40 56 100; do
  echo "-- B2TEX_TREE_SMEM_KB=$kb"; B2TEX_TREE_SMEM_KB=$kb timeout 300 python tools/run_pipeline.py C3 1 2>&1 | grep -E "rep 0|mrf\.k_tree" | tee -a gpurun_out/b_sweep.txt
done
for g in 8 16; do
  echo "-- B2TEX_MRF_GROUP=$g"; B2TEX_MRF_GROUP=$g timeout 300 python tools/run_pipeline.py C3 1 2>&1 | grep -E "rep 0|mrf\.k_tree" | tee -a gpurun_out/b_sweep.txt
done
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -c 6000 gpurun_out/b_bench.json; tail -5 gpurun_out/b_bench.err
