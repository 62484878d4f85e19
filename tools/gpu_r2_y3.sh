#!/usr/bin/env bash
set +e
mkdir -p gpurun_out
timeout 600 python tools/diag_n8192.py > gpurun_out/y3_diag.log 2>&1; tail -40 gpurun_out/y3_diag.log
