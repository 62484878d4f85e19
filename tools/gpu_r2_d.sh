#!/usr/bin/env bash
set +e
mkdir -p gpurun_out
for m in 0 1000000; do
  B200_MULTIBIT_LL_MAX=$m timeout 600 python tools/ab_bench.py --lib ours --what multibit3,multibit4 --batches 1,32,148,296,592,1184,4096 --steps 3 > gpurun_out/d_mb_ll$m.log 2>&1
  echo "LL max $m"; grep what gpurun_out/d_mb_ll$m.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:18],d['batch'],round(d['ms'],3))"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_bit or golden" > gpurun_out/d_pytest.log 2>&1; tail -4 gpurun_out/d_pytest.log
timeout 300 python tools/bench_mul.py --multi-bit > gpurun_out/d_mul_mb.log 2>&1; tail -1 gpurun_out/d_mul_mb.log
