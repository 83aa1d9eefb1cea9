"""Forward-NTT timings (device-resident, CUDA events) of the library selected by ZKB200_LIB: ms per transform at a few sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blockmaze_b200 as zk
from blockmaze_b200 import api
zk.init(0)
for logn, batch in ((16, 1), (18, 1), (18, 3), (20, 1), (24, 1)):
    ms = min(api.lib.zkb200_bench_ntt(logn, batch, 20) for _ in range(3))
    print("ntt 2^%d x%d: %.4f ms per transform" % (logn, batch, ms))
