#!/usr/bin/env bash
# round 2, GPU session B: v4 classic kernel vs v3, micro-benchmarks, multi-bit LL launch breakdown, tests
set +e
mkdir -p gpurun_out
B200_PBS_VARIANT=3 timeout 300 python tools/ab_bench.py --lib ours --what classic --batches 1,148,296,4096 --steps 3 > gpurun_out/b_classic_v3.log 2>&1
timeout 300 python tools/ab_bench.py --lib ours --what classic,kspbs --batches 1,148,296,4096 --steps 3 > gpurun_out/b_classic_v4.log 2>&1
tail -4 gpurun_out/b_classic_v3.log; tail -6 gpurun_out/b_classic_v4.log
(cd tools/micro && ./dsmem_bw > ../../gpurun_out/b_dsmem_bw.txt 2>&1; ./dfma_lat > ../../gpurun_out/b_dfma_lat.txt 2>&1)
cat gpurun_out/b_dsmem_bw.txt gpurun_out/b_dfma_lat.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_mb_launches.csv \
  python tools/ab_bench.py --lib ours --what multibit3,multibit4 --batches 1,148 --steps 1 > gpurun_out/b_ncu_mb.log 2>&1
python profiles/summarize.py launches gpurun_out/b_mb_launches.csv > gpurun_out/b_mb_launches.txt 2>&1; cat gpurun_out/b_mb_launches.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/b_pytest.log 2>&1; tail -8 gpurun_out/b_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_n2048_k1_l1_v3 -s 2 -c 1 -o /tmp/r2_v4 \
  python tools/ab_bench.py --lib ours --what classic --batches 4096 --steps 1 > gpurun_out/b_ncu_v4.log 2>&1
python profiles/summarize.py full /tmp/r2_v4.ncu-rep > gpurun_out/r2b_v4_pbs_full.txt 2>gpurun_out/b_sum.err; head -40 gpurun_out/r2b_v4_pbs_full.txt
du -sh gpurun_out
