#!/usr/bin/env bash
# round 2, GPU session Z4: fused inverse radix-16 pass (212 instead of 224 operations) in every kernel: tests + A/B numbers
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/z4_pytest.log 2>&1; tail -4 gpurun_out/z4_pytest.log
timeout 600 python tools/ab_bench.py --lib ours --what classic,kspbs,multibit3,multibit4 --batches 1,148,4096 --steps 4 > gpurun_out/z4_ab.log 2>&1
timeout 600 python tools/ab_bench.py --lib ours --what set11 --batches 148,4096 --steps 3 >> gpurun_out/z4_ab.log 2>&1
timeout 600 python tools/ab_bench.py --lib ours --what set33 --batches 148 --steps 2 >> gpurun_out/z4_ab.log 2>&1
grep what gpurun_out/z4_ab.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:28],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
