#!/usr/bin/env bash
# round 2, GPU session T: final validation -- full GPU suite incl. the cross-implementation parity test (reference .so present),
# smoke, bench with the reference_gpu block, launch list of the bench command, sanitizers on the new kernels
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/t_pytest.log 2>&1; tail -4 gpurun_out/t_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1; tail -1 gpurun_out/t_smoke.log
timeout 900 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; tail -c 400 gpurun_out/t_bench.json; echo; tail -2 gpurun_out/t_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/t_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/t_ncu_bench.log 2>&1
python profiles/summarize.py launches gpurun_out/t_launches.csv > gpurun_out/r2t_final_launches.txt 2>&1; cat gpurun_out/r2t_final_launches.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "zero_mask or n512_register_kernel_toy_sets or single_cmux or standalone" > gpurun_out/t_memcheck.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/t_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "zero_mask or single_cmux" > gpurun_out/t_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/t_racecheck.log | tail -4
