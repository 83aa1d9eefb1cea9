"""One forward NTT of 2^logn elements (default 24) for ncu captures of ntt_pass_kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blockmaze_b200 as zk
from blockmaze_b200 import api
zk.init(0)
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
print("ntt 2^%d: %.4f ms" % (logn, api.lib.zkb200_bench_ntt(logn, 1, 2)))
