#!/usr/bin/env bash
# round 2, GPU session C: classic kernel variants with warp-local exchange 2, multi-bit low-latency crossover, tests
set +e
mkdir -p gpurun_out
for v in 3 5 4; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 1,148,296,4096 --steps 4 > gpurun_out/c_classic_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/c_classic_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
for m in 0 1000000; do
  B200_MULTIBIT_LL_MAX=$m timeout 600 python tools/ab_bench.py --lib ours --what multibit3,multibit4 --batches 1,32,74,148,296,592 --steps 3 > gpurun_out/c_mb_ll$m.log 2>&1
  echo "LL max $m"; grep what gpurun_out/c_mb_ll$m.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:18],d['batch'],round(d['ms'],3))"
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c_pytest.log 2>&1; tail -6 gpurun_out/c_pytest.log
timeout 300 python tools/bench_mul.py > gpurun_out/c_mul.log 2>&1; tail -1 gpurun_out/c_mul.log
du -sh gpurun_out
