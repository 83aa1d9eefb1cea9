#!/bin/bash
# GPU run B: quad-cooperative MSM tails (parity first, with tight timeouts: a wrong shuffle mask hangs the kernel), then timing
mkdir -p gpurun_out/r02b
( timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 90 -k "msm_g1_edge or skewed" ) > gpurun_out/r02b/pytest0.log 2>&1
rc=$?; echo "pytest0 rc=$rc" >> gpurun_out/r02b/pytest0.log; tail -3 gpurun_out/r02b/pytest0.log
if [ $rc -ne 0 ]; then exit 1; fi
( timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prover.py -m gpu -x -q --timeout 120 -k "msm or proof_byte or cgo_genproof or synthetic or lanes" ) > gpurun_out/r02b/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b/pytest.log
tail -3 gpurun_out/r02b/pytest.log
for lib in ${LIBS:-libzkb200.so}; do
  ZKB200_LIB=$PWD/blockmaze_b200/$lib timeout 200 python scripts/gpu_quick.py send 2>&1 | grep QUICK >> gpurun_out/r02b/quick.jsonl
done
tail -n 4 gpurun_out/r02b/quick.jsonl
