#!/bin/bash
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
echo "== multi-rank on one GPU + MRF parity"
timeout 600 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/d_tests.txt
echo "== C3 pipeline, phase timers (G=16)"
B2TEX_MRF_GROUP=16 B2TEX_FOREST_TIMING=1 timeout 400 python tools/run_pipeline.py C3 2 2>&1 | grep -v "^trace" | tail -22 | tee gpurun_out/d_pipeline_c3.txt
for g in 8 32; do
  echo "-- B2TEX_MRF_GROUP=$g"; B2TEX_MRF_GROUP=$g timeout 300 python tools/run_pipeline.py C3 1 2>&1 | grep -E "rep 0|mrf\.k_tree" | tee -a gpurun_out/d_sweep.txt
done
for kb in 64 200; do
  echo "-- B2TEX_TREE_SMEM_KB=$kb (G=16)"; B2TEX_MRF_GROUP=16 B2TEX_TREE_SMEM_KB=$kb timeout 300 python tools/run_pipeline.py C3 1 2>&1 | grep -E "rep 0|mrf\.k_tree" | tee -a gpurun_out/d_sweep.txt
done
