#!/bin/bash
# multi-GPU: sharded check on C3s, then the bench with the seam phase timers
set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
cat mvs-texturing_b200/build/STAMP
for sc in ${CHECK_SCENES:-C3s}; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py $sc 2>&1 | grep -E "SHARDED|Error|error|Traceback" | head -8 | tee -a gpurun_out/mg_check_n$N.txt
done
B2TEX_SEAM_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$N" --steps 5 --warmup 3 ${BENCH_FLAGS:-} > gpurun_out/mg_bench_n$N.json 2> gpurun_out/mg_bench_n$N.err
grep -E "k_pcg_mg rank" gpurun_out/mg_bench_n$N.err | tail -$N
python - <<PY
import json
t=open('gpurun_out/mg_bench_n$N.json').read()
d=json.loads(t[t.find('{"metric'):].splitlines()[0])
print('N', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'stage', {k:round(v,2) for k,v in d['stage_ms'].items()})
for k in d['kernels'][:12]: print('  ', k['name'], round(k['ms_per_step'],2))
print('verify ok', d['verify'] and d['verify']['ok'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1))
PY
