#!/bin/bash
# round-2 late experiments (1 GPU): parity suite with the new defaults, then A/B of each switch on one send / deposit proof
#   ZKB200_SPMV_BITS (0/1 witness values skip the SpMV multiplication), ZKB200_FULL_POW (flat coset tables), ZKB200_SORT_V2 (absolute
#   cursors, all atomics of a scalar in flight, no memsets), ZKB200_FOLD_LANES (lanes per bucket in the piece fold of the H query),
#   ZKB200_NTT_TWSM (twiddles of a tile in shared memory)
OUT=gpurun_out/${1:-r03a}; mkdir -p $OUT; rm -f $OUT/quick.jsonl
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --durations=5 ) > $OUT/pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/pytest.log; tail -15 $OUT/pytest.log
ALLOFF="ZKB200_SPMV_BITS=0 ZKB200_FULL_POW=0 ZKB200_SORT_V2=0 ZKB200_FOLD_LANES=1 ZKB200_NTT_TWSM=0"
if [ $rc -ne 0 ]; then
  for v in "$ALLOFF" ZKB200_SPMV_BITS=0 ZKB200_FULL_POW=0 ZKB200_SORT_V2=0 ZKB200_FOLD_LANES=1 ZKB200_NTT_TWSM=0; do
    echo "== $v"; env $v timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prover.py -x -q --timeout 300 -k "not live" 2>&1 | tail -4
  done > $OUT/pytest_bisect.log 2>&1
  cat $OUT/pytest_bisect.log
fi
q() { # name circuit env...
  local name=$1 c=$2; shift 2
  env QUICK_SHORT=1 "$@" timeout 300 python scripts/gpu_quick.py $c 2>/dev/null | grep QUICK | sed "s/^QUICK /{\"variant\": \"$name\", \"circuit\": \"$c\", \"r\": /; s/$/}/" >> $OUT/quick.jsonl
}
q new send X=1
q alloff send $ALLOFF
q no_spmv_bits send ZKB200_SPMV_BITS=0
q no_full_pow send ZKB200_FULL_POW=0
q no_sort_v2 send ZKB200_SORT_V2=0
q fold_lanes1 send ZKB200_FOLD_LANES=1
q fold_lanes2 send ZKB200_FOLD_LANES=2
q no_twsm send ZKB200_NTT_TWSM=0
q new send X=1
q new deposit X=1
q alloff deposit $ALLOFF
q new mint X=1
q alloff mint $ALLOFF
cat $OUT/quick.jsonl
