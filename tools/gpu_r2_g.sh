#!/usr/bin/env bash
# round 2, GPU session G: v6 (tensor-memory exchange + one-slot TMA key ring) now fits two CTAs per SM; multi-bit TMA MAC fix
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "variants_are_bit_identical or multi_bit or golden" > gpurun_out/g_pytest.log 2>&1; tail -6 gpurun_out/g_pytest.log
for v in 9 11 10 12 13 8; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 1,148,296,592,4096 --steps 4 > gpurun_out/g_classic_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/g_classic_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
(cd tools/micro && for c in 1 296; do ./phase_clocks $c 6; done) > gpurun_out/g_phase_clocks.txt 2>&1; cat gpurun_out/g_phase_clocks.txt
B200_PBS_VARIANT=9 timeout 300 python tools/bench_mul.py > gpurun_out/g_mul_v9.log 2>&1; tail -1 gpurun_out/g_mul_v9.log
