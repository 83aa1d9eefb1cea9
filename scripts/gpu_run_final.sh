#!/bin/bash
# final evidence run (1 GPU): parity suite, default bench line, ncu launch lists (+FMA-pipe counts) of one send proof -- alone (team tails, two
# waves) and with the kernels a pipelined proof runs -- and ncu --set full of the accumulate kernels, the NTT passes of the QAP map and of a
# 2^24 transform, the H-query bucket reduction.  The ncu passes enqueue kernel by kernel (ZKB200_GRAPH=0): same kernels, fixed launch order.
OUT=gpurun_out/${1:-r03z}; mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --durations=5 ) > $OUT/pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
( time timeout 900 python bench.py --steps 200 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err; tail -c 200 $OUT/bench.json; echo
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/clocks_after_bench.csv
export ZKB200_GRAPH=0
M=gpu__time_duration.sum,sm__inst_executed_pipe_fma.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file $OUT/send_launches.csv python scripts/gpu_prove_once.py send 3 > $OUT/prove_once.log 2>&1
tail -1 $OUT/prove_once.log
ZKB200_TEAM=0 ZKB200_ACC_WAVES=1 timeout 600 ncu --metrics $M --clock-control none --csv --log-file $OUT/send_launches_pipelined.csv python scripts/gpu_prove_once.py send 3 > $OUT/prove_once_pipelined.log 2>&1
tail -1 $OUT/prove_once_pipelined.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -s 5 -c 5 -o $OUT/acc_full python scripts/gpu_prove_once.py send 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 10 -c 6 -o $OUT/ntt_qap_full python scripts/gpu_prove_once.py send 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 3 -c 3 -o $OUT/ntt24_full python scripts/gpu_ntt_once.py 24 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_team -s 5 -c 5 -o $OUT/reduce_team_full python scripts/gpu_prove_once.py send 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_kernel\|msm_scatter_abs -s 5 -c 5 -o $OUT/spmv_scatter_full python scripts/gpu_prove_once.py send 2 > /dev/null 2>&1
for f in acc_full ntt_qap_full ntt24_full reduce_team_full spmv_scatter_full; do ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/$f.raw.csv 2>/dev/null; done
rm -f $OUT/reduce_team_full.ncu-rep $OUT/spmv_scatter_full.ncu-rep      # (gpurun brings back at most 64 MiB: the raw pages of these two are enough)
ls -la $OUT
