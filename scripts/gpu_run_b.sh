#!/bin/bash
mkdir -p gpurun_out/r02k
( timeout 500 python -m pytest tests/test_gpu_prover.py tests/test_gpu_kernels.py -m gpu -x -q --timeout 200 -k "msm or proof_byte or synthetic or cgo_genproof" ) 2>&1 | tail -2
for c in send mint deposit; do
timeout 200 python scripts/gpu_quick.py $c 2>&1 | grep QUICK | sed "s/^QUICK {/QUICK {\"c\": \"$c\", /" >> gpurun_out/r02k/quick.jsonl
done
cut -c1-20,90-700 gpurun_out/r02k/quick.jsonl
