#!/bin/bash
# single GPU: TMA probe in its own process, all GPU tests, pipeline timers (patch stage split, k_tree prefetch on/off, 1/16-size
# scene), the bench line with e2e
set -u
mkdir -p gpurun_out
cat mvs-texturing_b200/build/STAMP
echo "== TMA probe (own process)"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k gradient_images 2>&1 | tail -4 | tee gpurun_out/r02_tma_probe.txt
if ! grep -q " passed" gpurun_out/r02_tma_probe.txt || grep -q failed gpurun_out/r02_tma_probe.txt; then
  echo "TMA kernel failed: continuing with B2TEX_TMA=0"; export B2TEX_TMA=0
fi
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_gpu_tests.txt
echo "== pipeline C3 (+ patch stages)"
timeout 400 python tools/run_pipeline.py C3 2 --patches 2>&1 | grep -v "^trace" | tail -32 | tee gpurun_out/r02_pipeline_c3.txt
echo "== B2TEX_TMA=0"
B2TEX_TMA=0 timeout 300 python tools/run_pipeline.py C3 2 2>&1 | grep -E "rep 1|k_lum_sobel" | tail -2
echo "== B2TEX_TREE_PREFETCH=0"
B2TEX_TREE_PREFETCH=0 timeout 300 python tools/run_pipeline.py C3 2 2>&1 | grep -E "rep 1|mrf.k_tree" | tail -2
echo "== C3s (1/16 size)"
timeout 300 python tools/run_pipeline.py C3s 3 2>&1 | grep -E "rep 2|mrf\.k_|k_pcg|seam_a" | tail -6
B2TEX_TREE_PREFETCH=0 timeout 300 python tools/run_pipeline.py C3s 3 2>&1 | grep -E "mrf.k_tree" | tail -1
echo "== bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python - <<PY
import json
t=open('gpurun_out/r02_bench_n1.json').read()
d=json.loads(t[t.find('{"metric'):].splitlines()[0])
print('ms', round(d['ms_per_step'],2), 'stage', {k:round(v,2) for k,v in d['stage_ms'].items()}, 'verify', d['verify']['ok'])
e=d['e2e']; print('e2e', round(e['ms_per_step'],1), 'pageable', e.get('pageable_host'), 'three', round(e['three_call_path']['ms_per_step'],1))
print('roofline', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
PY
tail -3 gpurun_out/r02_bench_n1.err
