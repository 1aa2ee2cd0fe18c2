#!/bin/bash
# final single-GPU pass of the round: TMA probe (own process), all GPU tests, the bench line, the ncu launch list of the bench
# command, ncu --set full captures of the top kernels, small-scene MRF timings (lanes per node)
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
echo "== TMA probe (own process)"
rm -f gpurun_out/r02_tma_probe.txt
for sc in tiny C2s; do timeout 200 python tools/tma_probe.py $sc 2>&1 | grep -E "TMA_PROBE" | head -6 | tee -a gpurun_out/r02_tma_probe.txt; done
if [ "$(grep -c 'wrong pixels: 0' gpurun_out/r02_tma_probe.txt)" != 2 ]; then echo "TMA kernel failed: continuing with B2TEX_TMA=0"; export B2TEX_TMA=0; fi
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests.txt
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python - <<PY
import json
t=open('gpurun_out/r02_bench_n1.json').read()
d=json.loads(t[t.find('{"metric'):].splitlines()[0])
print('ms', round(d['ms_per_step'],2), 'stage', {k:round(v,2) for k,v in d['stage_ms'].items()}, 'verify', d['verify']['ok'])
e=d['e2e']; print('e2e', round(e['ms_per_step'],1), 'pageable', round(e['pageable_host']['ms_per_step'],1), 'three', round(e['three_call_path']['ms_per_step'],1))
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
for k in d['kernels'][:10]: print('  ', k['name'], round(k['ms_per_step'],2), round(k['gbs'],1))
PY
tail -2 gpurun_out/r02_bench_n1.err
echo "== C3s: lanes per node"
for g in 16 8; do B2TEX_MRF_GROUP=$g timeout 200 python tools/run_pipeline.py C3s 3 2>&1 | grep -E "mrf.k_tree|mrf.k_forest" | tail -2; done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-verify --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv
for k in "k_tree<" k_lum_sobel; do
  name=$(echo $k | tr -d '<')
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$k" -s 2 -c 1 -f -o gpurun_out/r02_prof_$name python tools/run_pipeline.py C3 1 > gpurun_out/r02_ncu_$name.log 2>&1
  tail -1 gpurun_out/r02_ncu_$name.log | cut -c1-120
done
