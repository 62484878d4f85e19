#!/usr/bin/env bash
# round 2, GPU session E (first after the container was re-created): state of HEAD -- micro-benchmarks of the
# candidate pipes, classic variants incl. the TMA key ring, multi-bit low-latency incl. the TMA ring, full GPU tests, bench
set +e
mkdir -p gpurun_out
(cd tools/micro && timeout 120 ./pipes > ../../gpurun_out/e_pipes.txt 2>&1); cat gpurun_out/e_pipes.txt
for v in 5 7; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 1,74,148,296,4096 --steps 3 > gpurun_out/e_classic_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/e_classic_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
for t in 0 1; do
  B200_MULTIBIT_SEQ_TMA=$t timeout 600 python tools/ab_bench.py --lib ours --what multibit3,multibit4 --batches 1,32,148,296,4096 --steps 3 > gpurun_out/e_mb_tma$t.log 2>&1
  echo "multibit seq TMA $t"; grep what gpurun_out/e_mb_tma$t.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:18],d['batch'],round(d['ms'],3))"
done
timeout 300 python tools/bench_mul.py > gpurun_out/e_mul.log 2>&1; tail -1 gpurun_out/e_mul.log
timeout 300 python tools/bench_mul.py --multi-bit > gpurun_out/e_mul_mb.log 2>&1; tail -1 gpurun_out/e_mul_mb.log
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/e_pytest.log 2>&1; tail -6 gpurun_out/e_pytest.log
timeout 600 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; tail -c 3000 gpurun_out/e_bench.json
du -sh gpurun_out
