#!/usr/bin/env bash
# round 2, GPU session Z3: the library as committed last -- full GPU suite, smoke, bench
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/z3_pytest.log 2>&1; tail -4 gpurun_out/z3_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/z3_smoke.log 2>&1; tail -1 gpurun_out/z3_smoke.log
timeout 900 python bench.py > gpurun_out/z3_bench.json 2> gpurun_out/z3_bench.err; tail -c 300 gpurun_out/z3_bench.json; echo; tail -2 gpurun_out/z3_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z3_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['secondary']['frac'], d['clocks'])
for k,v in d['extras']['other_configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk!='config' and kk!='timing'})
PY
