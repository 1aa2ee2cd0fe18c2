#!/bin/bash
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
for mode in 1 2; do
  echo "== TMA probe mode $mode"
  B2TEX_TMA_MODE=$mode timeout 200 python tools/tma_probe.py tiny 2>&1 | grep -E "TMA_PROBE|Error|error" | head -8 | tee -a gpurun_out/r02_tma_probe.txt
done
export B2TEX_TMA=0
echo "== pytest -m gpu (TMA off)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_gpu_tests.txt
