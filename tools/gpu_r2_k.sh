#!/usr/bin/env bash
# round 2, GPU session K: one ncu --set full capture of EVERY kernel of the library (tools/ncu_all_kernels.py)
set +e
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none -k 'regex:^(bsk_|centered_|forward_fft|glwe_|keyswitch|ks_|mb_|modulus_|pbs_|seeded_)' -o /tmp/r2k_all python tools/ncu_all_kernels.py > gpurun_out/k_ncu_all.log 2>&1; tail -3 gpurun_out/k_ncu_all.log
python profiles/summarize.py table /tmp/r2k_all.ncu-rep > gpurun_out/r2k_all_kernels.txt 2> gpurun_out/k_sum.err; grep -c "^kernel" gpurun_out/r2k_all_kernels.txt; head -30 gpurun_out/r2k_all_kernels.txt
du -sh gpurun_out
