#!/bin/bash
# k_tree variants on C3: lanes per node x resident CTAs per SM; MRF parity tests first; one ncu capture of the default
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi_rank.py -m gpu -x -q 2>&1 | tail -3
for cfg in "16 3" "16 2" "8 3" "8 2" "32 2"; do
  set -- $cfg
  echo "== G=$1 blocks=$2"
  B2TEX_MRF_GROUP=$1 B2TEX_TREE_BLOCKS=$2 timeout 300 python tools/run_pipeline.py C3 2 2>&1 | grep -E "rep 1|mrf\.k_tree|mrf\.k_forest" | tail -3 | tee -a gpurun_out/r02_tree_variants.txt
done
echo "== ncu k_tree (3rd launch)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tree -s 5 -c 1 -f -o gpurun_out/r02_prof_k_tree python tools/run_pipeline.py C3 1 > gpurun_out/r02_ncu_k_tree.log 2>&1
tail -2 gpurun_out/r02_ncu_k_tree.log | cut -c1-200
