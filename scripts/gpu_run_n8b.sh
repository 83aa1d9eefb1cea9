#!/bin/bash
# one process, 8 GPUs, proofs replayed from CUDA graphs: 4096 and 1024 mixed transactions
N=${N:-8}; OUT=gpurun_out/${1:-r03p}; mkdir -p $OUT
( time timeout 200 python bench.py --workload mixed1024 --single-process --gpus $N --repeat 4 ) > $OUT/sp_x4_graph.json 2> $OUT/sp_x4_graph.err
echo "x4 rc=$?"; cut -c1-420 $OUT/sp_x4_graph.json
( time timeout 200 python bench.py --workload mixed1024 --single-process --gpus $N ) > $OUT/sp_1024_graph.json 2> $OUT/sp_1024_graph.err
echo "1024 rc=$?"; cut -c1-420 $OUT/sp_1024_graph.json
