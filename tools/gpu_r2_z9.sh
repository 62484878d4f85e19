#!/usr/bin/env bash
# round 2, GPU session Z9: phase offset between the two resident CTAs of an SM (B200_P22_STAGGER sweep), batch 4096 and 296
set +e
mkdir -p gpurun_out
for st in 0 1000 2000 3000 4000 5000 6000 8000; do
B200_P22_STAGGER=$st timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 296,4096 --steps 4 > gpurun_out/z9_ab_$st.log 2>&1
echo "stagger $st"; grep what gpurun_out/z9_ab_$st.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['ms_best'],3),round(d.get('pbs_per_s',0)))"; done
