#!/bin/bash
# last hardware pass of round 2: all GPU tests on the final build, the patch stages with the seam planning on the device,
# then the C5 workload (BASELINE.json configs[4]: 10 M faces / 1000 views at 4K) on ONE GPU
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests.txt
echo "== C3 pipeline + patch stages"
timeout 200 python tools/run_pipeline.py C3 2 --patches 2>&1 | grep -v "^trace" | grep -E "rep 1|patches|tp\.|ls\.|local_seam|texture_patches|k_lum_sobel|mrf.k_tree" | tail -16 | tee gpurun_out/r02_pipeline_c3_patches.txt
echo "== C5 on one GPU"
timeout 330 python bench.py --workload C5 --steps 2 --warmup 1 --no-verify --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c5_n1.json 2> gpurun_out/r02_bench_c5_n1.err
tail -c 2500 gpurun_out/r02_bench_c5_n1.json; tail -4 gpurun_out/r02_bench_c5_n1.err | cut -c1-300
