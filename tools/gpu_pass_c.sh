#!/bin/bash
# round-2 GPU pass C (1 GPU): CTA-cooperative k_tree, block-aggregated frontier pushes, multi-rank test
set -u
mkdir -p gpurun_out
echo "== gpu tests (MRF parity + multi-rank on one GPU)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi_rank.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/c_tests.txt
echo "== C3 pipeline, forest phase timers"
B2TEX_FOREST_TIMING=1 timeout 400 python tools/run_pipeline.py C3 2 2>&1 | grep -v "^trace" | tail -22 | tee gpurun_out/c_pipeline_c3.txt
echo "== smem / group sweep"
for kb in 64 150 200; do
  echo "-- B2TEX_TREE_SMEM_KB=$kb"; B2TEX_TREE_SMEM_KB=$kb timeout 300 python tools/run_pipeline.py C3 1 2>&1 | grep -E "rep 0|mrf\.k_tree" | tee -a gpurun_out/c_sweep.txt
done
for g in 8 16; do
  echo "-- B2TEX_MRF_GROUP=$g"; B2TEX_MRF_GROUP=$g timeout 300 python tools/run_pipeline.py C3 1 2>&1 | grep -E "rep 0|mrf\.k_tree" | tee -a gpurun_out/c_sweep.txt
done
