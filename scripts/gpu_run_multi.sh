#!/bin/bash
# multi-GPU validation: N = number of GPUs on the box
N=${N:-2}
mkdir -p gpurun_out/r02m
nvidia-smi -L > gpurun_out/r02m/gpus_n$N.txt
( timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q --timeout 200 -k "prove_batch_spreads" ) > gpurun_out/r02m/pytest_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02m/pytest_n$N.log; tail -3 gpurun_out/r02m/pytest_n$N.log
( time timeout 600 python bench.py --workload mixed1024 --single-process --gpus $N ) > gpurun_out/r02m/mixed_single_process_n$N.json 2> gpurun_out/r02m/mixed_single_process_n$N.err
echo "single-process rc=$?"; cut -c1-400 gpurun_out/r02m/mixed_single_process_n$N.json
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 3 ) > gpurun_out/r02m/bench_n$N.json 2> gpurun_out/r02m/bench_n$N.err
echo "torchrun rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02m/bench_n$N.json").read().strip().splitlines()[-1])
    print("value",d["value"],"e2e",d["e2e"]["value"],"mixed",d["mixed1024"]["value"],"split g1 ms",d["msm_split24"]["g1"]["ms_per_msm"],d["msm_split24"]["g1"]["result_matches_golden"],"g2 ms",d["msm_split24"]["g2"]["ms_per_msm"],d["msm_split24"]["g2"]["result_matches_golden"])
except Exception as e:
    print("parse failed",e)
PY
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) > gpurun_out/r02m/ref_n$N.json 2> gpurun_out/r02m/ref_n$N.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r02m/ref_n$N.json
