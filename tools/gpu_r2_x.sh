#!/usr/bin/env bash
# round 2, GPU session X: shipped state after the split post-MAC barrier: full suite, bench, ncu --set full of the shipped kernel
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/x_pytest.log 2>&1; tail -3 gpurun_out/x_pytest.log
timeout 900 python bench.py > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; tail -c 300 gpurun_out/x_bench.json; echo; tail -2 gpurun_out/x_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_n2048_k1_l1_v -s 2 -c 1 -o /tmp/r2x_ship python tools/ab_bench.py --lib ours --what classic --batches 4096 --steps 1 > gpurun_out/x_ncu.log 2>&1
python profiles/summarize.py full /tmp/r2x_ship.ncu-rep > gpurun_out/r2x_shipped_pbs_full.txt 2> gpurun_out/x_sum.err; head -40 gpurun_out/r2x_shipped_pbs_full.txt
timeout 300 python tools/ab_bench.py --lib ours --what classic,kspbs --batches 1,32,148,296,592,4096 --steps 4 > gpurun_out/x_classic.log 2>&1
grep what gpurun_out/x_classic.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:20],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
