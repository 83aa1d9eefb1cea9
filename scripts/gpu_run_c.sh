#!/bin/bash
mkdir -p gpurun_out/r02c
timeout 400 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02c/send_launches.csv python scripts/gpu_prove_once.py send 3 > gpurun_out/r02c/prove_once.log 2>&1
tail -2 gpurun_out/r02c/prove_once.log
