#!/usr/bin/env bash
# round 2, GPU session H: ncu --set full of the v6 hybrid (variant 11), v6 ring (9) and v7 x3 (13) kernels, batch 4096
set +e
mkdir -p gpurun_out
for v in 11 13 9; do
  B200_PBS_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on -k regex:pbs_n2048_k1_l1_v -s 2 -c 1 -o /tmp/r2h_v$v \
    python tools/ab_bench.py --lib ours --what classic --batches 4096 --steps 1 > gpurun_out/h_ncu_v$v.log 2>&1
  python profiles/summarize.py full /tmp/r2h_v$v.ncu-rep > gpurun_out/r2h_v${v}_pbs_full.txt 2> gpurun_out/h_sum_v$v.err
  head -45 gpurun_out/r2h_v${v}_pbs_full.txt
done
du -sh gpurun_out
