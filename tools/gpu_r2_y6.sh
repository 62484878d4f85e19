#!/usr/bin/env bash
# round 2, GPU session Y6: N = 8192 kernel as committed (l = 2 digit specialisation): tests, A/B against the reference library, ncu
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "large_polynomial or n8192" > gpurun_out/y6_pytest.log 2>&1; tail -3 gpurun_out/y6_pytest.log
timeout 400 python tools/ab_bench.py --lib ours --what set33 --batches 1,32,148,296,592,1024 --steps 2 --out gpurun_out/r2y6_ab_ours.json > gpurun_out/y6_ab.log 2>&1
timeout 600 python tools/ab_bench.py --lib ref --what set33 --batches 1,32,148 --steps 1 --out gpurun_out/r2y6_ab_ref.json > gpurun_out/y6_ab_ref.log 2>&1
for f in gpurun_out/y6_ab.log gpurun_out/y6_ab_ref.log; do grep what $f | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['lib'],d['what'][:20],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_n8192 -c 1 -o /tmp/r2y6_n8192 python tools/ab_bench.py --lib ours --what set33 --batches 148 --steps 1 > gpurun_out/y6_ncu.log 2>&1
python profiles/summarize.py full /tmp/r2y6_n8192.ncu-rep > gpurun_out/r2y6_n8192_full.txt 2> gpurun_out/y6_sum.err; sed -n 1,12p gpurun_out/r2y6_n8192_full.txt
cp /tmp/r2y6_n8192.ncu-rep gpurun_out/ 2>/dev/null
