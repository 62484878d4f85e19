#!/usr/bin/env bash
# round 2, GPU session Z: final validation with the N = 8192 tensor-memory kernel in the library -- full GPU suite, smoke,
# bench (with the reference_gpu block and the 3_3 extras line), launch list of the bench command, sanitizers on the new kernel,
# one ncu --set full capture of every kernel
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; tail -4 gpurun_out/z_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/z_smoke.log 2>&1; tail -1 gpurun_out/z_smoke.log
timeout 900 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; tail -c 600 gpurun_out/z_bench.json; echo; tail -2 gpurun_out/z_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/z_ncu_bench.log 2>&1
python profiles/summarize.py launches gpurun_out/z_launches.csv > gpurun_out/r2z_final_launches.txt 2>&1; cat gpurun_out/r2z_final_launches.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_polynomial or n8192_register_kernel_matches" > gpurun_out/z_memcheck.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/z_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_polynomial" > gpurun_out/z_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/z_racecheck.log | tail -4
timeout 1500 ncu --set full --clock-control none -k 'regex:^(bsk_|centered_|forward_fft|glwe_|keyswitch|ks_|mb_|modulus_|pbs_|seeded_)' -o /tmp/r2z_all python tools/ncu_all_kernels.py > gpurun_out/z_ncu_all.log 2>&1
python profiles/summarize.py table /tmp/r2z_all.ncu-rep > gpurun_out/r2z_all_kernels.txt 2> gpurun_out/z_table.err; wc -l gpurun_out/r2z_all_kernels.txt
