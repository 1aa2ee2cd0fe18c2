#!/bin/bash
# one `ncu --set full` capture per hot kernel of the C3 pipeline (third launch where the kernel repeats) + phase timers
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
KERNELS=${KERNELS:-"k_tree k_forest k_rays k_quality k_pcg k_lum_sobel"}
for k in $KERNELS; do
  skip=2; case $k in k_rays|k_quality|k_pcg|k_lum_sobel|k_cull) skip=0;; esac
  echo "== ncu $k (skip $skip)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/r02_prof_$k \
      python tools/run_pipeline.py ${SCENE:-C3} 1 > gpurun_out/r02_ncu_$k.log 2>&1
  tail -2 gpurun_out/r02_ncu_$k.log | cut -c1-300
done
echo "== phase timers (B2TEX_FOREST_TIMING)"
B2TEX_FOREST_TIMING=1 timeout 300 python tools/run_pipeline.py ${SCENE:-C3} 2 2>&1 | grep -E "k_forest phases|k_tree|rep 1" | tee gpurun_out/r02_forest_timing.txt
ls -la gpurun_out/*.ncu-rep
