#!/bin/bash
mkdir -p gpurun_out/r02c
ZKB200_AFFINE_ALWAYS=1 timeout 400 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02c/send_launches_aff.csv python scripts/gpu_prove_once.py send 3 > gpurun_out/r02c/prove_once.log 2>&1
tail -2 gpurun_out/r02c/prove_once.log
