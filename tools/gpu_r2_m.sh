#!/usr/bin/env bash
# round 2, GPU session M: rotate + decompose trims on the v6 kernels: digits converted with an exponent splice + DADD instead of I2F (bit 0), sign as a predicate (bit 1)
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants_are_bit_identical" > gpurun_out/m_pytest.log 2>&1; tail -3 gpurun_out/m_pytest.log
for v in 14 15; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 1,148 --steps 5 > gpurun_out/m_classic_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/m_classic_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
for v in 11 16 17 18; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 296,4096 --steps 5 > gpurun_out/m_classic_big_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/m_classic_big_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
