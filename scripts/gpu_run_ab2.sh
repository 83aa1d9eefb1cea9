#!/bin/bash
# occupancy of the accumulate kernel against the number of proofs in flight (pipelined throughput), waves of a proof alone; new defaults per circuit
OUT=gpurun_out/${1:-r03b}; mkdir -p $OUT; rm -f $OUT/quick.jsonl
q() { # name circuit env...
  local name=$1 c=$2; shift 2
  env QUICK_SHORT=1 "$@" timeout 300 python scripts/gpu_quick.py $c 2>/dev/null | grep QUICK | sed "s/^QUICK /{\"variant\": \"$name\", \"circuit\": \"$c\", \"r\": /; s/$/}/" >> $OUT/quick.jsonl
}
q new send X=1
for ctas in 2 3; do for lanes in 3 4 5; do q "acc${ctas}_lanes${lanes}" send ZKB200_ACC_CTAS=$ctas ZKB200_LANES=$lanes; done; done
q lanes4 send ZKB200_LANES=4
q waves1 send ZKB200_ACC_WAVES=1
q waves3 send ZKB200_ACC_WAVES=3
q new deposit X=1
q new mint X=1
q new redeem X=1
q acc3_lanes4 deposit ZKB200_ACC_CTAS=3 ZKB200_LANES=4
q acc3_lanes4 mint ZKB200_ACC_CTAS=3 ZKB200_LANES=4
cat $OUT/quick.jsonl
