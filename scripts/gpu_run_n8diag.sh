#!/bin/bash
# why does ONE process feed 8 GPUs slower over a long batch than over a short one?  single process (spin / blocking waits, 3 / 4 callers per GPU)
# against one process per GPU on the same 4096 transactions
N=${N:-8}; OUT=gpurun_out/${1:-r03o}; mkdir -p $OUT
run() { # name env... -- args
  local name=$1; shift
  ( time env "$@" timeout 300 python bench.py --workload mixed1024 --single-process --gpus $N --repeat 4 ) > $OUT/sp_$name.json 2> $OUT/sp_$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.loads(open('$OUT/sp_$name.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['seconds'],3), d['proofs_per_device'])" 2>&1)"
}
run spin X=1
run block ZKB200_SPIN=0
run callers4 BENCH_CALLERS=4
run callers2 BENCH_CALLERS=2
( time timeout 300 python bench.py --workload mixed1024 --single-process --gpus $N ) > $OUT/sp_1024.json 2> $OUT/sp_1024.err
echo "1024 rc=$?"; cut -c1-330 $OUT/sp_1024.json
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --workload mixed1024 --gpus $N --repeat 4 ) > $OUT/mp_x4.json 2> $OUT/mp_x4.err
echo "torchrun x4 rc=$?"; cut -c1-330 $OUT/mp_x4.json
nproc; grep -c processor /proc/cpuinfo
