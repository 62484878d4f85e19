#!/usr/bin/env bash
# round 2, GPU session Z7: end state (fused inverse pass everywhere; accumulator read-back kept in the N = 2048 kernels, dropped in N = 512):
# full GPU suite, smoke, bench, launch list
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/z7_pytest.log 2>&1; tail -3 gpurun_out/z7_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/z7_smoke.log 2>&1; tail -1 gpurun_out/z7_smoke.log
timeout 900 python bench.py > gpurun_out/z7_bench.json 2> gpurun_out/z7_bench.err; tail -2 gpurun_out/z7_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z7_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['secondary']['frac'], d['clocks'], d['reference_gpu']['pbs_per_s'])
print(d['extras']['ks_pbs_per_s_this_rank'])
for k,v in d['extras']['other_configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk!='config' and kk!='timing'})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z7_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/z7_ncu_bench.log 2>&1
python profiles/summarize.py launches gpurun_out/z7_launches.csv > gpurun_out/r2z7_final_launches.txt 2>&1; head -8 gpurun_out/r2z7_final_launches.txt
timeout 300 python tools/ab_bench.py --lib ours --what classic,kspbs --batches 1,148,4096 --steps 4 > gpurun_out/z7_ab.log 2>&1
timeout 300 python tools/ab_bench.py --lib ours --what set11 --batches 148,4096 --steps 3 >> gpurun_out/z7_ab.log 2>&1
grep what gpurun_out/z7_ab.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:28],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
