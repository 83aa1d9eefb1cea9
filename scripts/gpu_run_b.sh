#!/bin/bash
# GPU run: parity subset with tight timeouts, then quick timing (LIBS = library variants to time)
mkdir -p gpurun_out/r02b
( timeout 500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prover.py -m gpu -x -q --timeout 120 -k "${TESTS:-field or domain or ntt or msm or proof_byte or cgo_genproof or synthetic or lanes or qap}" ) > gpurun_out/r02b/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02b/pytest.log
tail -3 gpurun_out/r02b/pytest.log
for lib in ${LIBS:-libzkb200.so}; do
  ZKB200_LIB=$PWD/blockmaze_b200/$lib timeout 200 python scripts/gpu_quick.py send 2>&1 | grep QUICK >> gpurun_out/r02b/quick.jsonl
done
tail -n 4 gpurun_out/r02b/quick.jsonl
