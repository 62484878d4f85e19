#!/usr/bin/env bash
# round 2, GPU session Z11: bench.py twice (per-step times in the line)
set +e
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/z11_bench_$i.json 2> gpurun_out/z11_bench_$i.err
python - <<PY
import json
d=json.loads(open('gpurun_out/z11_bench_$i.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['step_ms'], d['clocks']['sm_mhz'], d['clocks']['samples'])
PY
done
