#!/usr/bin/env bash
# round 2, GPU session Y: first run of the N = 8192 tensor-memory kernel (3_3): tests, A/B against the workspace kernel, ncu
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_polynomial or n8192" > gpurun_out/y_pytest.log 2>&1; tail -15 gpurun_out/y_pytest.log
timeout 400 python tools/ab_bench.py --lib ours --what set33 --batches 1,148,296,1024 --steps 2 --out gpurun_out/y_ab_tmem.json > gpurun_out/y_ab_tmem.log 2>&1
B200_N8192_GENERIC=1 timeout 400 python tools/ab_bench.py --lib ours --what set33 --batches 1,148 --steps 2 --out gpurun_out/y_ab_ws.json > gpurun_out/y_ab_ws.log 2>&1
for f in gpurun_out/y_ab_tmem.log gpurun_out/y_ab_ws.log; do grep what $f | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:20],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"; tail -2 $f | grep -v what; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_n8192 -c 1 -o /tmp/r2y_n8192 python tools/ab_bench.py --lib ours --what set33 --batches 148 --steps 1 > gpurun_out/y_ncu.log 2>&1
python profiles/summarize.py full /tmp/r2y_n8192.ncu-rep > gpurun_out/r2y_n8192_full.txt 2> gpurun_out/y_sum.err; head -60 gpurun_out/r2y_n8192_full.txt
cp /tmp/r2y_n8192.ncu-rep gpurun_out/ 2>/dev/null
