import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockmaze_b200 as zk
zk.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262143
for c in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "12,-12,-14,-16,-17".split(","))]:
    print("msm g1 n=%d c=%d: %.3f ms" % (n, c, zk.lib.zkb200_bench_msm(1, n, c, 3)), flush=True)
