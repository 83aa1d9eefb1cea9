"""Golden hashes for the key generator: runs the reference's <circuit>_key tools (oracle/_ref) under LD_PRELOAD=libfixed_rng.so
with ZK_FIXED_SEED pinned and records size + sha256 of the pk/vk files they write.  Build container only (35-110 s per circuit)."""
import hashlib, json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
SEED = 7
out = {}
procs = {}
for c in ("mint", "send", "deposit", "redeem"):
    d = os.path.join(REF, "tmp_keygen", c)
    os.makedirs(d, exist_ok=True)
    if not os.path.exists(os.path.join(d, c + "vk.txt")) or "--force" in sys.argv:
        env = dict(os.environ, LD_PRELOAD=os.path.join(REF, "libfixed_rng.so"), ZK_FIXED_SEED=str(SEED))
        procs[c] = subprocess.Popen([os.path.join(REF, c + "_key")], cwd=d, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
for c, p in procs.items():
    p.wait()
for c in ("mint", "send", "deposit", "redeem"):
    d = os.path.join(REF, "tmp_keygen", c)
    pk, vk = open(os.path.join(d, c + "pk.txt"), "rb").read(), open(os.path.join(d, c + "vk.txt"), "rb").read()
    out[c] = dict(seed=SEED, pk_size=len(pk), vk_size=len(vk), pk_sha256=hashlib.sha256(pk).hexdigest(), vk_sha256=hashlib.sha256(vk).hexdigest())
json.dump(out, open(os.path.join(HERE, "keygen.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
