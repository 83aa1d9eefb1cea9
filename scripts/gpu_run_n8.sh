#!/bin/bash
# one process, N GPUs (N = GPUs of the box): the library's own device scheduler behind zkb200_prove_batch / gen*proof
N=${N:-8}; OUT=gpurun_out/${1:-r03n}; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_n$N.txt
( timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q --timeout 250 -k "prove_batch_spreads" ) > $OUT/pytest_n$N.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_n$N.log; tail -3 $OUT/pytest_n$N.log
( time timeout 400 python bench.py --workload mixed1024 --single-process --gpus $N ) > $OUT/mixed_single_process_n$N.json 2> $OUT/mixed_single_process_n$N.err
echo "single-process rc=$?"; cut -c1-500 $OUT/mixed_single_process_n$N.json
( time timeout 400 python bench.py --workload mixed1024 --single-process --gpus $N --repeat 4 ) > $OUT/mixed_single_process_x4_n$N.json 2> $OUT/mixed_single_process_x4_n$N.err
echo "single-process x4 rc=$?"; cut -c1-500 $OUT/mixed_single_process_x4_n$N.json
