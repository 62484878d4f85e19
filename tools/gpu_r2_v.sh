#!/usr/bin/env bash
# round 2, GPU session V: v6 without the CTA barrier after the MAC (split arrive / wait, variants 20 / 21); serde ingest on the GPU
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "variants_are_bit_identical or serialized_keys" > gpurun_out/v_pytest.log 2>&1; tail -3 gpurun_out/v_pytest.log
for v in 17 20; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 296,592,4096 --steps 5 > gpurun_out/v_classic_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/v_classic_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
for v in 19 21; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 1,148 --steps 5 > gpurun_out/v_classic_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/v_classic_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
B200_PBS_VARIANT=20 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "zero_mask or single_cmux" > gpurun_out/v_racecheck20.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/v_racecheck20.log | tail -2
