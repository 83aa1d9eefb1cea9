#!/bin/bash
mkdir -p gpurun_out/r02d
( ZKB200_AFFINE_ALWAYS=1 timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q --timeout 150 -k "proof_byte or synthetic" ) > gpurun_out/r02d/pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/r02d/pytest.log
tail -3 gpurun_out/r02d/pytest.log
for cfg in "2 2 1" "2 4 1" "1 4 1"; do
  set -- $cfg
  ZKB200_AFFINE_ROUNDS=$1 ZKB200_AFF_CTAS=$2 ZKB200_AFFINE_ALWAYS=$3 timeout 200 python scripts/gpu_quick.py send 2>&1 | grep QUICK | sed "s/^QUICK {/QUICK {\"rounds\": $1, \"ctas\": $2, \"always\": $3, /" >> gpurun_out/r02d/quick3.jsonl
done
cat gpurun_out/r02d/quick3.jsonl
ZKB200_AFFINE_ALWAYS=1 ZKB200_AFF_CTAS=4 timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum,smsp__inst_executed.sum -k regex:affine --clock-control none --csv --log-file gpurun_out/r02d/aff_inl.csv python scripts/gpu_prove_once.py send 2 > /dev/null 2>&1
grep affine gpurun_out/r02d/aff_inl.csv | tail -6 | cut -c1-60,150-400
