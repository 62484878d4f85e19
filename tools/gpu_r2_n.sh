#!/usr/bin/env bash
# round 2, GPU session N: register kernel for N = 512 (PARAM_MESSAGE_1_CARRY_1): tests, then timing against the generic kernel
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "n512 or generic_kernel or seeded" > gpurun_out/n_pytest.log 2>&1; tail -15 gpurun_out/n_pytest.log
timeout 600 python tools/ab_bench.py --lib ours --what set11 --batches 1,148,296,1024,4096 --steps 3 > gpurun_out/n_set11_reg.log 2>&1
echo "register kernel"; grep what gpurun_out/n_set11_reg.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
tail -2 gpurun_out/n_set11_reg.log
B200_N512_GENERIC=1 timeout 600 python tools/ab_bench.py --lib ours --what set11 --batches 148,1024 --steps 3 > gpurun_out/n_set11_gen.log 2>&1
echo "generic kernel"; grep what gpurun_out/n_set11_gen.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
