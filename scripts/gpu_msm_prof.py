import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import blockmaze_b200 as zk
zk.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 196607
c = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = int(sys.argv[3]) if len(sys.argv) > 3 else 1
print("msm g%d n=%d c=%d: %.3f ms" % (g, n, c, zk.lib.zkb200_bench_msm(g, n, c, 2)))
