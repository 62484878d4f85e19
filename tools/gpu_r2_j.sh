#!/usr/bin/env bash
# round 2, GPU session J: new default dispatch (v6 ring <= one CTA per SM, v6 hybrid above): full GPU suite, bench, FheUint64 mul
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/j_pytest.log 2>&1; tail -6 gpurun_out/j_pytest.log
timeout 300 python tools/ab_bench.py --lib ours --what classic,kspbs --batches 1,32,148,296,592,4096 --steps 4 > gpurun_out/j_classic.log 2>&1
grep what gpurun_out/j_classic.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:20],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
timeout 300 python tools/bench_mul.py > gpurun_out/j_mul.log 2>&1; tail -1 gpurun_out/j_mul.log
timeout 300 python tools/bench_mul.py --multi-bit > gpurun_out/j_mul_mb.log 2>&1; tail -1 gpurun_out/j_mul_mb.log
timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; tail -c 1500 gpurun_out/j_bench.json; tail -3 gpurun_out/j_bench.err
