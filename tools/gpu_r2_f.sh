#!/usr/bin/env bash
# round 2, GPU session F: exchange 2 through tensor memory (variant 8) -- bit identity, then timing against variant 5
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants_are_bit_identical or zero_mask or single_cmux" > gpurun_out/f_pytest.log 2>&1; tail -15 gpurun_out/f_pytest.log
for v in 5 8 10 11 9 12 13; do
  B200_PBS_VARIANT=$v timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 1,148,296,4096 --steps 4 > gpurun_out/f_classic_var$v.log 2>&1
  echo "variant $v"; grep what gpurun_out/f_classic_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['batch'],round(d['ms'],3),round(d['pbs_per_s']))"
done
(cd tools/micro && for c in 1 296; do ./phase_clocks $c; ./phase_clocks $c 6; done) > gpurun_out/f_phase_clocks.txt 2>&1; cat gpurun_out/f_phase_clocks.txt
B200_PBS_VARIANT=12 timeout 300 python tools/bench_mul.py > gpurun_out/f_mul_v12.log 2>&1; tail -1 gpurun_out/f_mul_v12.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_bit or golden" > gpurun_out/f_pytest_mb.log 2>&1; tail -3 gpurun_out/f_pytest_mb.log
