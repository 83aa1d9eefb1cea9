"""Quick per-variant measurement on the GPU box (select the library with ZKB200_LIB): NTT batch timing at circuit size, one send
proof alone (CUDA-event phases), pipelined throughput with three proofs in flight, single-caller genSendproof latency."""
import json, os, statistics, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blockmaze_b200 as zk
from blockmaze_b200 import api, wallet as F
zk.init(0)
kd = os.path.join(ROOT, "oracle/_ref/prfKey")
api.set_key_dir(kd)
out = {"lib": os.path.basename(api.LIB_PATH)}
out["ntt18x3_ms"] = round(api.lib.zkb200_bench_ntt(18, 3, 50) * 3, 4)
out["ntt24_ms"] = round(api.lib.zkb200_bench_ntt(24, 1, 5), 3)
circuit = sys.argv[1] if len(sys.argv) > 1 else "send"
pk = zk.ProvingKey(os.path.join(kd, circuit + "pk.txt"))
tx = F.synthetic(circuit, 1)
w = api.witness(circuit, tx)
r, s = 12345, 67890
res = pk.prove(w, r, s)
assert res["rc"] == 0 and api.verify_proof(circuit, res["proof_hex"], api.verify_args(circuit, tx))
ph = [[] for _ in range(8)]
for i in range(25):
    api.lib.zkb200_flush_l2()
    res = pk.prove(None, r, s)
    if i >= 5:
        for k in range(8):
            ph[k].append(res["timings_ms"][k])
names = ["gpu", "qap", "msm_h", "host_tail", "acc_h", "a_done", "b_done", "l_done"]
out["alone_ms"] = {n: round(statistics.mean(v), 3) for n, v in zip(names, ph)}
depth = pk.lanes
lanes = [pk.lane_acquire() for _ in range(depth)]
for ln in lanes:
    pk.submit(ln, w, r, s)
for ln in lanes:
    pk.collect(ln)
steps = 300
api.lib.zkb200_device_timer(0)
for i in range(steps):
    ln = lanes[i % depth]
    if i >= depth:
        pk.collect(ln)
    pk.submit(ln, None, r, s)
for i in range(steps, steps + depth):
    pk.collect(lanes[i % depth])
ms = float(api.lib.zkb200_device_timer(1))
for ln in lanes:
    pk.lane_release(ln)
out["pipelined_ms_per_proof"] = round(ms / steps, 4)
if os.environ.get("QUICK_SHORT"):
    print("QUICK " + json.dumps(out)); sys.exit(0)
lat, brk = [], []
for i in range(25):
    t = time.perf_counter(); api.gen_proof(circuit, F.synthetic(circuit, 100 + i)); lat.append(1e3 * (time.perf_counter() - t)); brk.append(api.last_breakdown_ms())
out["p50_latency_ms"] = round(statistics.median(lat[5:]), 3)
out["breakdown_ms"] = {k: round(statistics.median(b[k] for b in brk[5:]), 3) for k in brk[0]}
from concurrent.futures import ThreadPoolExecutor
txs = [F.synthetic(circuit, 1000 + i) for i in range(300)]
with ThreadPoolExecutor(depth) as pool:
    list(pool.map(lambda tx: api.gen_proof(circuit, tx), txs[:12]))
    t = time.perf_counter(); list(pool.map(lambda tx: api.gen_proof(circuit, tx), txs)); out["e2e_proofs_per_s"] = round(len(txs) / (time.perf_counter() - t), 1)
print("QUICK " + json.dumps(out))
