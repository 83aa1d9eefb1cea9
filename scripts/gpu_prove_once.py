"""Load one circuit's pk, prove the golden assignment a few times (for ncu launch lists / full captures)."""
import json, os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blockmaze_b200 as zk
c = sys.argv[1] if len(sys.argv) > 1 else "send"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
zk.init(0)
g = json.load(open(os.path.join(ROOT, "tests/golden/%s.json" % c)))
w = zlib.decompress(open(os.path.join(ROOT, "tests/golden/%s_assignment.bin.z" % c), "rb").read())
pk = zk.ProvingKey(os.path.join(ROOT, "oracle/_ref/prfKey", c + "pk.txt"))
for i in range(reps):
    res = pk.prove(w if i == 0 else None, int(g["r"], 16), int(g["s"], 16))
    print(c, "equal" if res["proof_hex"] == g["proof_hex"] else "MISMATCH", res["timings_ms"], res["launches"])
