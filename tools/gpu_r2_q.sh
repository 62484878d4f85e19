#!/usr/bin/env bash
# round 2, GPU session Q: full GPU suite with the N = 512 kernel in place, 1_1 timing (automatic mode), ncu of the ring kernel
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/q_pytest.log 2>&1; tail -4 gpurun_out/q_pytest.log
timeout 600 python tools/ab_bench.py --lib ours --what set11 --batches 1,148,296,444,1024,4096 --steps 3 > gpurun_out/q_set11.log 2>&1
grep what gpurun_out/q_set11.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
timeout 600 ncu --set full --clock-control none -k regex:pbs_n512 -s 1 -c 1 -o /tmp/r2q_n512 python tools/ab_bench.py --lib ours --what set11 --batches 444 --steps 1 > gpurun_out/q_ncu.log 2>&1
python profiles/summarize.py full /tmp/r2q_n512.ncu-rep > gpurun_out/r2q_n512_ring_full.txt 2> gpurun_out/q_sum.err; head -36 gpurun_out/r2q_n512_ring_full.txt
