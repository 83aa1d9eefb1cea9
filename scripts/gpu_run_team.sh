#!/bin/bash
# team tails (msm_team.cuh) against the one-thread-per-chain tails: microbenchmark, parity suite, quick A/B per circuit
OUT=gpurun_out/r02p; mkdir -p $OUT; rm -f $OUT/quick.jsonl
timeout 120 ./scripts/ubench/team > $OUT/team.jsonl 2>&1; cat $OUT/team.jsonl
if grep -q "illegal\|failure" $OUT/team.jsonl; then timeout 300 compute-sanitizer --tool memcheck --print-limit 3 ./scripts/ubench/team 2>&1 | head -80 > $OUT/team_sanitizer.log; head -40 $OUT/team_sanitizer.log; fi
if [ -n "$TEAM_DEBUG" ]; then
  for m in 1 2 4 7; do
    echo "== ZKB200_TEAM=$m"; ZKB200_TEAM=$m timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "msm" --timeout 300 2>&1 | tail -8
  done
  exit 0
fi
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --durations=5 ) > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -12 $OUT/pytest.log
for c in send mint deposit; do
  timeout 300 python scripts/gpu_quick.py $c 2>/dev/null | grep QUICK | sed "s/^QUICK /{\"variant\": \"team\", \"circuit\": \"$c\", \"r\": /; s/$/}/" >> $OUT/quick.jsonl
  ZKB200_TEAM=0 timeout 300 python scripts/gpu_quick.py $c 2>/dev/null | grep QUICK | sed "s/^QUICK /{\"variant\": \"old\", \"circuit\": \"$c\", \"r\": /; s/$/}/" >> $OUT/quick.jsonl
done
cat $OUT/quick.jsonl
