#!/bin/bash
OUT=gpurun_out/${1:-r03i}; mkdir -p $OUT
( timeout 300 python bench.py --workload mixed1024 --single-process --gpus 1 ) > $OUT/mixed_sp1.json 2> $OUT/mixed_sp1.err
echo "mixed rc=$?"; grep -n "zkb200: CUDA" $OUT/mixed_sp1.err | head -3; cut -c1-200 $OUT/mixed_sp1.json
( timeout 300 python -m pytest tests/test_gpu_prover.py -m gpu -x -q --timeout 250 -k "prove_batch_spreads or concurrent or overlapping or malformed" ) > $OUT/tsel.log 2>&1
echo "tests rc=$?"; tail -3 $OUT/tsel.log
