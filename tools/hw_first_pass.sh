#!/bin/bash
# First GPU call of the next round: everything written after the round-1 GPU budget ran out, on hardware, in one go.
#   gpurun --timeout 1500 -- 'bash tools/hw_first_pass.sh'            (1 GPU)
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/hw_first_pass.sh p2p'  (2 GPUs: the fused multi-GPU seam kernel)
set -u
mkdir -p gpurun_out
if [ "${1:-}" = "p2p" ]; then
    N=$(nvidia-smi -L | wc -l)
    echo "== sharded check, replicated seam solve ($N GPUs)"
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py 2>&1 | tail -5
    echo "== sharded check, B2TEX_SEAM_P2P=1 (csrc/seam_mg.cu)"
    B2TEX_SEAM_P2P=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 tools/check_sharded.py 2>&1 | tail -5
    exit 0
fi
echo "== late round-1 GPU tests (never run on hardware before)"
timeout 900 python -m pytest tests/test_zz_gpu_occlusion.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/zz_tests.txt
echo "== C3 pipeline incl. texture patches + local seam leveling, per-kernel events"
timeout 500 python tools/run_pipeline.py C3 2 --patches 2>&1 | tail -60 | tee gpurun_out/pipeline_c3_patches.txt
echo "== bench with the extra stages"
timeout 400 python bench.py --steps 3 --warmup 3 --with-patches --no-cpu-baseline > gpurun_out/bench_with_patches.json 2> gpurun_out/bench_with_patches.err
tail -c 1500 gpurun_out/bench_with_patches.json
