#!/usr/bin/env bash
# round 2, GPU session O: N = 512 register kernel with the persistent grid + 3-chunk key ring: tests, timing, ncu
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "n512" > gpurun_out/o_pytest.log 2>&1; tail -5 gpurun_out/o_pytest.log; grep -i "cuda error\|panic" gpurun_out/o_pytest.log | head -3
timeout 600 python tools/ab_bench.py --lib ours --what set11 --batches 1,148,296,592,1024,4096 --steps 3 > gpurun_out/o_set11_reg.log 2>&1
echo "register kernel"; grep what gpurun_out/o_set11_reg.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
timeout 600 ncu --set full --clock-control none -k regex:pbs_n512 -s 1 -c 1 -o /tmp/r2o_n512 python tools/ab_bench.py --lib ours --what set11 --batches 592 --steps 1 > gpurun_out/o_ncu.log 2>&1
python profiles/summarize.py full /tmp/r2o_n512.ncu-rep > gpurun_out/r2o_n512_full.txt 2> gpurun_out/o_sum.err; head -50 gpurun_out/r2o_n512_full.txt
