#!/bin/bash
# GPU run A: whole parity suite (tight per-test timeout), default bench line, launch list with FMA-pipe counts of one send proof
mkdir -p gpurun_out/r02e
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --durations=8 ) > gpurun_out/r02e/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e/pytest.log
tail -4 gpurun_out/r02e/pytest.log
if [ "${SKIP_BENCH:-0}" = "0" ]; then
( time timeout 900 python bench.py --steps 200 --warmup 3 ) > gpurun_out/r02e/bench.json 2> gpurun_out/r02e/bench.err
echo "bench rc=$?" >> gpurun_out/r02e/bench.err
tail -c 300 gpurun_out/r02e/bench.json
fi
