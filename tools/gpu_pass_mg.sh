#!/bin/bash
# multi-GPU pass: sharded pipeline (peer-memory MRF exchange + fused seam PCG) vs the oracle, then the bench
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
cat mvs-texturing_b200/build/STAMP
for sc in C1d C3s; do
  echo "== sharded check $sc on $N GPUs"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py $sc 2>&1 | grep -E "SHARDED|Error|error|Traceback" | head -8 | tee -a gpurun_out/mg_check_n$N.txt
done
echo "== bench on $N GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$N" --steps 5 --warmup 3 ${BENCH_FLAGS:-} > gpurun_out/mg_bench_n$N.json 2> gpurun_out/mg_bench_n$N.err
tail -c 3500 gpurun_out/mg_bench_n$N.json; tail -5 gpurun_out/mg_bench_n$N.err
