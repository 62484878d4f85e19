#!/usr/bin/env bash
# round 2, GPU session A: tests, same-box A/B against the reference's CUDA backend, bench, ncu of the reference kernel
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
timeout 600 python tools/ab_bench.py --lib ours --out gpurun_out/ab_ours.json > gpurun_out/a_ab_ours.log 2>&1; echo "ab ours rc=$?"
timeout 900 python tools/ab_bench.py --lib ref --out gpurun_out/ab_ref.json > gpurun_out/a_ab_ref.log 2>&1; echo "ab ref rc=$?"
tail -3 gpurun_out/a_ab_ref.log
timeout 900 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_ref_launches.csv \
  python tools/ab_bench.py --lib ref --what classic,ks,multibit4 --batches 1,4096 --steps 1 > gpurun_out/a_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:specialized_2_2 -s 2 -c 1 -o /tmp/r2_ref_pbs_2_2 \
  python tools/ab_bench.py --lib ref --what classic --batches 4096 --steps 1 > gpurun_out/a_ncu2.log 2>&1; echo "ncu full rc=$?"
# the report embeds the 60 MB fat binary: keep the text pages, not the .ncu-rep (gpurun_out is capped at 64 MiB)
ncu -i /tmp/r2_ref_pbs_2_2.ncu-rep --page raw --csv > gpurun_out/r2_ref_pbs_2_2_raw.csv 2>/dev/null
ncu -i /tmp/r2_ref_pbs_2_2.ncu-rep --page details > gpurun_out/r2_ref_pbs_2_2_details.txt 2>/dev/null
python profiles/summarize.py full /tmp/r2_ref_pbs_2_2.ncu-rep > gpurun_out/r2_ref_pbs_2_2_full.txt 2>gpurun_out/a_sum.err
ls -la gpurun_out | tail -20
