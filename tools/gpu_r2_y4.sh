#!/usr/bin/env bash
# round 2, GPU session Y4: N = 8192 kernel: ties-even noise check, start stagger sweep
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "large_polynomial or n8192" > gpurun_out/y4_pytest.log 2>&1; tail -5 gpurun_out/y4_pytest.log
for st in 0 15000 30000 60000 120000 240000; do
B200_N8192_STAGGER=$st timeout 300 python tools/ab_bench.py --lib ours --what set33 --batches 148,296 --steps 2 > gpurun_out/y4_ab_$st.log 2>&1
echo "stagger $st"; grep what gpurun_out/y4_ab_$st.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:20],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"; done
