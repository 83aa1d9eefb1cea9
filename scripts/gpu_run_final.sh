#!/bin/bash
# final evidence run (1 GPU): parity suite, default bench line, ncu launch list (+FMA-pipe counts) of one send proof,
# ncu --set full of the H-query accumulate kernel, the NTT passes of the QAP map and of a 2^24 transform
OUT=gpurun_out/r02y; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --durations=5 ) > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
( time timeout 900 python bench.py --steps 200 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err; tail -c 200 $OUT/bench.json; echo
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/clocks_after_bench.csv
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/send_launches.csv python scripts/gpu_prove_once.py send 3 > $OUT/prove_once.log 2>&1
tail -1 $OUT/prove_once.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 6 -c 2 -o $OUT/acc_full python scripts/gpu_prove_once.py send 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 10 -c 6 -o $OUT/ntt_qap_full python scripts/gpu_prove_once.py send 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 3 -c 3 -o $OUT/ntt24_full python scripts/gpu_ntt_once.py 24 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_team -s 3 -c 1 -o $OUT/reduce_team_h_full python scripts/gpu_prove_once.py send 1 > /dev/null 2>&1
ls -la $OUT
