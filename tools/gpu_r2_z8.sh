#!/usr/bin/env bash
# round 2, GPU session Z8 (2 GPUs): the torchrun path of bench.py with the end-state library, both arms
set +e
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/z8_bench2.json 2> gpurun_out/z8_bench2.err; tail -c 900 gpurun_out/z8_bench2.json | head -c 900; echo; tail -3 gpurun_out/z8_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/z8_ref2.json 2> gpurun_out/z8_ref2.err; tail -c 600 gpurun_out/z8_ref2.json; echo; tail -2 gpurun_out/z8_ref2.err
