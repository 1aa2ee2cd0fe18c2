#!/bin/bash
# round-2 GPU pass A (1 GPU): new MRF kernels (k_forest tree layout + k_tree) on hardware
set -u
mkdir -p gpurun_out
echo "== memcheck of the smoke pipeline (small scene)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke 2>&1 | tail -15 | tee gpurun_out/a_memcheck.txt
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/a_tests.txt
echo "== C3 pipeline incl. texture patches + local seam leveling, per-kernel events"
timeout 500 python tools/run_pipeline.py C3 2 --patches 2>&1 | tail -70 | tee gpurun_out/a_pipeline_c3.txt
echo "== smem sweep"
for kb in 48 100; do
  echo "-- B2TEX_TREE_SMEM_KB=$kb"; B2TEX_TREE_SMEM_KB=$kb timeout 300 python tools/run_pipeline.py C3 1 2>&1 | grep -E "rep 0|mrf\." | tee -a gpurun_out/a_smem_sweep.txt
done
