#!/usr/bin/env bash
# round 2, GPU session Z5: shipped state with the fused inverse radix-16 pass: ncu --set full of the headline kernel, bench, launch list
set +e
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_n2048_k1_l1_v -s 2 -c 1 -o /tmp/r2z5_ship python tools/ab_bench.py --lib ours --what classic --batches 4096 --steps 1 > gpurun_out/z5_ncu.log 2>&1
python profiles/summarize.py full /tmp/r2z5_ship.ncu-rep > gpurun_out/r2z5_shipped_pbs_full.txt 2> gpurun_out/z5_sum.err; head -48 gpurun_out/r2z5_shipped_pbs_full.txt
timeout 900 python bench.py > gpurun_out/z5_bench.json 2> gpurun_out/z5_bench.err; tail -2 gpurun_out/z5_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/z5_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['secondary']['frac'], d['clocks'], d['reference_gpu']['pbs_per_s'])
for k,v in d['extras']['other_configs'].items(): print(k, {kk:vv for kk,vv in v.items() if kk!='config' and kk!='timing'})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z5_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/z5_ncu_bench.log 2>&1
python profiles/summarize.py launches gpurun_out/z5_launches.csv > gpurun_out/r2z5_final_launches.txt 2>&1; head -8 gpurun_out/r2z5_final_launches.txt
