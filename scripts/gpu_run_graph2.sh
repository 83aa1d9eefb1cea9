#!/bin/bash
OUT=gpurun_out/${1:-r03h}; mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -s --timeout 250 -k "prove_batch_spreads" ) > $OUT/t73.log 2>&1
echo "t73 rc=$?"; grep -n "zkb200\|CUDA\|Abort\|passed\|failed" $OUT/t73.log | head -20
( timeout 300 python bench.py --workload mixed1024 --single-process --gpus 1 ) > $OUT/mixed_sp1.json 2> $OUT/mixed_sp1.err
echo "mixed rc=$?"; grep -n "zkb200: CUDA" $OUT/mixed_sp1.err | head; cut -c1-200 $OUT/mixed_sp1.json
( ZKB200_GRAPH=0 timeout 300 python bench.py --workload mixed1024 --single-process --gpus 1 ) > $OUT/mixed_sp1_nograph.json 2> $OUT/mixed_sp1_nograph.err
echo "mixed nograph rc=$?"; cut -c1-200 $OUT/mixed_sp1_nograph.json
