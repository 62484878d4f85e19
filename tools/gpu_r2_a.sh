#!/usr/bin/env bash
# round 2, GPU session A: tests, same-box A/B against the reference's CUDA backend, bench, ncu of the reference kernel
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python tools/ab_bench.py --lib ours --out gpurun_out/ab_ours.json > gpurun_out/a_ab_ours.log 2>&1; echo "ab ours rc=$?"
timeout 900 python tools/ab_bench.py --lib ref --out gpurun_out/ab_ref.json > gpurun_out/a_ab_ref.log 2>&1; echo "ab ref rc=$?"
tail -3 gpurun_out/a_ab_ref.log
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ref_launches.csv \
  python tools/ab_bench.py --lib ref --what classic,ks,multibit4 --batches 1,4096 --steps 1 > gpurun_out/a_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:specialized_2_2 -s 2 -c 1 -o gpurun_out/r2_ref_pbs_2_2 \
  python tools/ab_bench.py --lib ref --what classic --batches 4096 --steps 1 > gpurun_out/a_ncu2.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
