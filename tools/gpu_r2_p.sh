#!/usr/bin/env bash
# round 2, GPU session P: N = 512 register kernel with the TMA key ring (3 or 2 LWEs per CTA) against the register-ring mode
set +e
mkdir -p gpurun_out
for m in 0 2 1; do
  B200_N512_MODE=$m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "n512" > gpurun_out/p_pytest_m$m.log 2>&1; echo "mode $m: $(tail -1 gpurun_out/p_pytest_m$m.log)"
done
for m in 0 2; do
  B200_N512_MODE=$m timeout 600 python tools/ab_bench.py --lib ours --what set11 --batches 1,148,444,1024,4096 --steps 3 > gpurun_out/p_set11_m$m.log 2>&1
  echo "mode $m"; grep what gpurun_out/p_set11_m$m.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
done
