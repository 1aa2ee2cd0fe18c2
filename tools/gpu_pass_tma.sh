#!/bin/bash
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
rm -f gpurun_out/r02_tma_probe.txt
for sc in tiny C2s occ2 C3s; do timeout 200 python tools/tma_probe.py $sc 2>&1 | grep -E "TMA_PROBE" | head -4 | tee -a gpurun_out/r02_tma_probe.txt; done
if [ "$(grep -c 'wrong pixels: 0' gpurun_out/r02_tma_probe.txt)" != 4 ]; then echo "TMA kernel failed"; exit 0; fi
echo "== C3 pipeline with the TMA kernel"
timeout 300 python tools/run_pipeline.py C3 2 2>&1 | grep -E "rep 1|k_lum_sobel|mrf.k_tree" | tail -3 | tee gpurun_out/r02_tma_c3.txt
echo "== gradient + data-cost + multi-rank tests"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi_rank.py -m gpu -x -q 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_lum_sobel -c 1 -f -o gpurun_out/r02_prof_k_lum_sobel python tools/run_pipeline.py C3 1 > gpurun_out/r02_ncu_k_lum_sobel.log 2>&1
tail -1 gpurun_out/r02_ncu_k_lum_sobel.log | cut -c1-100
