#!/usr/bin/env bash
# round 2, GPU session L: the reference's CUDA backend on the other classic sets (and P22 again) on the same box
set +e
mkdir -p gpurun_out
timeout 900 python tools/ab_bench.py --lib ref --what set11,set33 --batches 148,1024 --steps 3 > gpurun_out/l_sets_ref.log 2>&1
grep -E "what|unavailable" gpurun_out/l_sets_ref.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d.get('what','?')[:40],d.get('batch'),round(d.get('ms',0),3),round(d.get('pbs_per_s',0)))"
tail -3 gpurun_out/l_sets_ref.log
timeout 900 python tools/ab_bench.py --lib ours --what set11,set33 --batches 148,1024 --steps 3 > gpurun_out/l_sets_ours.log 2>&1
grep what gpurun_out/l_sets_ours.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:40],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
