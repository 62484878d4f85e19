#!/usr/bin/env bash
# round 2, GPU session S: N = 512 kernel: padded accumulator rows + own spectrum from registers in the MAC
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "n512" > gpurun_out/s_pytest.log 2>&1; tail -2 gpurun_out/s_pytest.log
timeout 600 python tools/ab_bench.py --lib ours --what set11 --batches 1,148,296,444,4096 --steps 3 > gpurun_out/s_set11.log 2>&1
grep what gpurun_out/s_set11.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
