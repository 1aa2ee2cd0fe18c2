#!/bin/bash
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
export B2TEX_MRF_GROUP=16
echo "== ncu k_tree (3rd launch)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tree -s 2 -c 1 -f -o gpurun_out/prof_k_tree python tools/run_pipeline.py C3 1 > gpurun_out/f_ncu_tree.log 2>&1
tail -3 gpurun_out/f_ncu_tree.log
echo "== ncu k_forest (3rd launch)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_forest -s 2 -c 1 -f -o gpurun_out/prof_k_forest python tools/run_pipeline.py C3 1 > gpurun_out/f_ncu_forest.log 2>&1
tail -3 gpurun_out/f_ncu_forest.log
ls -la gpurun_out/*.ncu-rep
