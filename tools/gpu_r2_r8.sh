#!/usr/bin/env bash
# round 2, 8-GPU session: weak scaling (4096 per GPU), BASELINE configs[4] (65,536 over 8 GPUs), strong scaling (4096 total)
set +e
mkdir -p gpurun_out
run() { # name, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 --steps 5 --warmup 3 ${@:3} > gpurun_out/r8_$2.json 2> gpurun_out/r8_$2.err
  tail -c 600 gpurun_out/r8_$2.json; echo; tail -2 gpurun_out/r8_$2.err
}
run 29511 weak
run 29512 config4_65536 --batch-global 65536
run 29513 strong --scaling strong
nvidia-smi topo -m > gpurun_out/r8_topo.txt 2>&1
