#!/bin/bash
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
echo "== N=1, multi-GPU seam kernel on one rank"
B2TEX_SEAM_MG1=1 B2TEX_SEAM_TIMING=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-verify > gpurun_out/mg1_bench.json 2> gpurun_out/mg1_bench.err
grep -E "k_pcg_mg rank" gpurun_out/mg1_bench.err | tail -1
python - <<PY
import json
t=open('gpurun_out/mg1_bench.json').read()
d=json.loads(t[t.find('{"metric'):].splitlines()[0])
print('N', d['n_gpus'], 'ms', round(d['ms_per_step'],2), 'stage', {k:round(v,2) for k,v in d['stage_ms'].items()}, [ (k['name'], round(k['ms_per_step'],2)) for k in d['kernels'] if 'pcg' in k['name']])
PY
bash tools/gpu_pass_mg2.sh
