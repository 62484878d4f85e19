#!/usr/bin/env bash
# round 2, GPU session W: v6 hybrid with both CTA barriers of a step split (variant 22) against variant 20
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "variants_are_bit_identical" > gpurun_out/w_pytest.log 2>&1; tail -2 gpurun_out/w_pytest.log
for v in 20 22; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 296,592,4096 --steps 5 > gpurun_out/w_classic_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/w_classic_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
B200_PBS_VARIANT=22 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "zero_mask or single_cmux" > gpurun_out/w_racecheck22.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/w_racecheck22.log | tail -2
