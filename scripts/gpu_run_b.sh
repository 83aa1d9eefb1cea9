#!/bin/bash
OUT=gpurun_out/r02z; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 9 -c 1 -o $OUT/acc_h_full python scripts/gpu_prove_once.py send 2 > /dev/null 2>&1
cat > /tmp/one_gen.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import blockmaze_b200 as zk
from blockmaze_b200 import api, wallet as F
zk.init(0); api.set_key_dir(os.path.join(os.getcwd(), "oracle/_ref/prfKey"))
api.gen_proof("send", F.synthetic("send", 1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sha256_witness -c 1 -o $OUT/sha_full python /tmp/one_gen.py > /dev/null 2>&1
ls -la $OUT/*.ncu-rep
