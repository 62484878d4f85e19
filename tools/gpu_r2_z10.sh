#!/usr/bin/env bash
# round 2, GPU session Z10: last validation of the library as committed (full GPU suite, smoke, bench)
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/z10_pytest.log 2>&1; tail -3 gpurun_out/z10_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/z10_smoke.log 2>&1; tail -1 gpurun_out/z10_smoke.log
timeout 900 python bench.py > gpurun_out/z10_bench.json 2> gpurun_out/z10_bench.err; tail -2 gpurun_out/z10_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z10_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['secondary']['frac'], d['clocks'], d['reference_gpu']['pbs_per_s'], d['gpu_launches'])
for k,v in d['extras']['other_configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk!='config' and kk!='timing'})
PY
