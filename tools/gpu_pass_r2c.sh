#!/bin/bash
# single GPU: all GPU tests, then pipeline timers (patch stage split, TMA Sobel vs vector Sobel, k_tree prefetch on/off)
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_gpu_tests.txt
echo "== default (TMA Sobel, prefetch)"
timeout 400 python tools/run_pipeline.py C3 2 --patches 2>&1 | grep -v "^trace" | tail -28 | tee gpurun_out/r02_pipeline_c3.txt
echo "== B2TEX_NO_TMA=1"
B2TEX_NO_TMA=1 timeout 300 python tools/run_pipeline.py C3 2 2>&1 | grep -E "rep 1|k_lum_sobel" | tail -2
echo "== B2TEX_TREE_PREFETCH=0"
B2TEX_TREE_PREFETCH=0 timeout 300 python tools/run_pipeline.py C3 2 2>&1 | grep -E "rep 1|mrf.k_tree" | tail -2
echo "== C3s (1/16 size: the per-rank problem of an 8-GPU run, roughly)"
timeout 300 python tools/run_pipeline.py C3s 3 2>&1 | grep -E "rep 2|mrf\.|k_pcg|seam" | tail -8
B2TEX_TREE_PREFETCH=0 timeout 300 python tools/run_pipeline.py C3s 3 2>&1 | grep -E "rep 2|mrf.k_tree" | tail -2
