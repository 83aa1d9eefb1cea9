#!/bin/bash
# CUDA graph per lane + completion signals in pinned memory: parity suite, then graph on/off per circuit
OUT=gpurun_out/${1:-r03g}; mkdir -p $OUT; rm -f $OUT/quick.jsonl
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --durations=5 ) > $OUT/pytest.log 2>&1
rc=$?; echo "pytest rc=$rc" >> $OUT/pytest.log; tail -15 $OUT/pytest.log
if [ $rc -ne 0 ]; then
  ( ZKB200_GRAPH=0 timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 ) > $OUT/pytest_nograph.log 2>&1
  echo "pytest (no graph) rc=$?" >> $OUT/pytest_nograph.log; tail -8 $OUT/pytest_nograph.log
fi
q() { # name circuit env...
  local name=$1 c=$2; shift 2
  env "$@" timeout 300 python scripts/gpu_quick.py $c 2>$OUT/quick_$name_$c.err | grep QUICK | sed "s/^QUICK /{\"variant\": \"$name\", \"circuit\": \"$c\", \"r\": /; s/$/}/" >> $OUT/quick.jsonl
}
q graph send X=1
q nograph send ZKB200_GRAPH=0
q graph deposit QUICK_SHORT=1
q nograph deposit QUICK_SHORT=1 ZKB200_GRAPH=0
q graph mint QUICK_SHORT=1
cat $OUT/quick.jsonl
