#!/usr/bin/env bash
# round 2, GPU session Y5: N = 8192 kernel with twiddles requested one pass ahead
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "large_polynomial or n8192" > gpurun_out/y5_pytest.log 2>&1; tail -3 gpurun_out/y5_pytest.log
timeout 300 python tools/ab_bench.py --lib ours --what set33 --batches 1,148,296 --steps 2 > gpurun_out/y5_ab.log 2>&1
grep what gpurun_out/y5_ab.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:20],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_n8192 -c 1 -o /tmp/r2y5_n8192 python tools/ab_bench.py --lib ours --what set33 --batches 148 --steps 1 > gpurun_out/y5_ncu.log 2>&1
python profiles/summarize.py full /tmp/r2y5_n8192.ncu-rep > gpurun_out/r2y5_n8192_full.txt 2> gpurun_out/y5_sum.err; sed -n 1,36p gpurun_out/r2y5_n8192_full.txt
cp /tmp/r2y5_n8192.ncu-rep gpurun_out/ 2>/dev/null
