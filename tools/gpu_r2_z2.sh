#!/usr/bin/env bash
# round 2, GPU session Z2: racecheck of the N = 8192 kernel's debug instance (every thread arrives on the ring's empty barriers)
set +e
mkdir -p gpurun_out
B200_N8192_RACECHECK=1 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_polynomial and tmem" > gpurun_out/z2_racecheck_all_arrive.log 2>&1; grep -E "RACECHECK SUMMARY|Error: Race|passed|failed" gpurun_out/z2_racecheck_all_arrive.log | cut -c1-200 | tail -6
B200_N8192_GEN1=1 timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large_polynomial and tmem" > gpurun_out/z2_racecheck_gen1.log 2>&1; grep -E "RACECHECK SUMMARY|Error: Race|passed|failed" gpurun_out/z2_racecheck_gen1.log | cut -c1-200 | tail -6
