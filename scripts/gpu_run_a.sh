#!/bin/bash
# GPU run A (round 2): parity suite, default bench line, launch list with FMA-pipe counts of one send proof
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a/smi.txt 2>&1
nproc >> gpurun_out/r02a/smi.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/r02a/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a/pytest.log
( time timeout 900 python bench.py --steps 200 --warmup 3 ) > gpurun_out/r02a/bench.json 2> gpurun_out/r02a/bench.err
echo "bench rc=$?" >> gpurun_out/r02a/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum --clock-control none --csv --log-file gpurun_out/r02a/send_launches.csv python scripts/gpu_prove_once.py send 3 > gpurun_out/r02a/prove_once.log 2>&1
tail -3 gpurun_out/r02a/pytest.log
tail -c 600 gpurun_out/r02a/bench.json
