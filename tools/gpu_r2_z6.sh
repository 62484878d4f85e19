#!/usr/bin/env bash
# round 2, GPU session Z6: accumulator update without the redundant shared-memory read; source-level ncu of the headline kernel
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants_are_bit_identical or zero_mask or n512 or single_cmux or p22_ks_pbs" > gpurun_out/z6_pytest.log 2>&1; tail -3 gpurun_out/z6_pytest.log
timeout 600 python tools/ab_bench.py --lib ours --what classic,kspbs --batches 1,148,4096 --steps 4 > gpurun_out/z6_ab.log 2>&1
timeout 600 python tools/ab_bench.py --lib ours --what set11 --batches 148,4096 --steps 3 >> gpurun_out/z6_ab.log 2>&1
grep what gpurun_out/z6_ab.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  ',d['what'][:28],d['batch'],round(d['ms'],3),round(d.get('pbs_per_s',0)))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_n2048_k1_l1_v -s 2 -c 1 -o /tmp/r2z6_ship python tools/ab_bench.py --lib ours --what classic --batches 592 --steps 1 > gpurun_out/z6_ncu.log 2>&1
cp /tmp/r2z6_ship.ncu-rep gpurun_out/ 2>/dev/null
python profiles/summarize.py full /tmp/r2z6_ship.ncu-rep 2>/dev/null | sed -n 1,16p
