#!/bin/bash
mkdir -p gpurun_out/r02f
( timeout 600 python -m pytest tests/test_gpu_prover.py tests/test_cabi_link.py -m gpu -x -q --timeout 200 -k "not live_reference and not keygen" ) > gpurun_out/r02f/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f/pytest.log
tail -4 gpurun_out/r02f/pytest.log
for c in send deposit mint; do for gw in 1 0; do
  ZKB200_GPU_WITNESS=$gw timeout 200 python scripts/gpu_quick.py $c 2>&1 | grep QUICK | sed "s/^QUICK {/QUICK {\"gpu_witness\": $gw, \"circuit\": \"$c\", /" >> gpurun_out/r02f/quick2.jsonl
done; done
cut -c1-60,230-600 gpurun_out/r02f/quick2.jsonl
