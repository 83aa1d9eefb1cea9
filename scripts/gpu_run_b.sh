#!/bin/bash
mkdir -p gpurun_out/r02h
( timeout 600 python -m pytest tests/test_gpu_prover.py tests/test_gpu_kernels.py -m gpu -x -q --timeout 200 -k "not live_reference and not keygen and not 2_24" ) > gpurun_out/r02h/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02h/pytest.log
tail -3 gpurun_out/r02h/pytest.log
for seven in 0 1; do
ZKB200_QAP_SEVEN=$seven timeout 200 python scripts/gpu_quick.py send 2>&1 | grep QUICK | sed "s/^QUICK {/QUICK {\"seven\": $seven, /" >> gpurun_out/r02h/quick.jsonl
done
ZKB200_QAP_SEVEN=0 timeout 200 python scripts/gpu_quick.py deposit 2>&1 | grep QUICK | sed "s/^QUICK {/QUICK {\"seven\": 0, \"c\": \"deposit\", /" >> gpurun_out/r02h/quick.jsonl
cut -c1-40,100-700 gpurun_out/r02h/quick.jsonl
