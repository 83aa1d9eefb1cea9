"""First GPU bring-up: field layer, NTT domains, MSM, full prover vs golden.  Run under gpurun."""
import json, os, random, sys, time, zlib, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blockmaze_b200 as zk
from oracle import bn254_oracle as O
from oracle import refapi as Rf
zk.init(0)
rng = random.Random(7)
def fb(v): return b"".join(int(x).to_bytes(32, "little") for x in v)
def fl(b): return [int.from_bytes(b[i:i+32], "little") for i in range(0, len(b), 32)]

# 1. field layer (raw Montgomery representatives)
for name, p in (("fr", O.R_MOD), ("fq", O.Q_MOD)):
    a = [0, 1, p - 1, p - 2, (1 << 253)] + [rng.randrange(p) for _ in range(5000)]
    b = [p - 1, 0, p - 1, 1, 12345] + [rng.randrange(p) for _ in range(5000)]
    Ri = pow(1 << 256, -1, p)
    got = fl(zk.field_op(name, "mul", fb(a), fb(b)))
    assert got == [x * y * Ri % p for x, y in zip(a, b)], name + " mul"
    assert fl(zk.field_op(name, "add", fb(a), fb(b))) == [(x + y) % p for x, y in zip(a, b)], name + " add"
    assert fl(zk.field_op(name, "sub", fb(a), fb(b))) == [(x - y) % p for x, y in zip(a, b)], name + " sub"
    assert fl(zk.field_op(name, "to_mont", fb(a))) == [x * (1 << 256) % p for x in a]
    assert fl(zk.field_op(name, "from_mont", fb(a))) == [x * Ri % p for x in a]
    inv = fl(zk.field_op(name, "inverse", fb(a[:64])))
    assert inv == [(pow(x * Ri % p, -1, p) * (1 << 256) % p if x else 0) for x in a[:64]], name + " inv"
print("field ok", flush=True)

# 2. domains vs python oracle (small) and vs reference (large)
for ms in (2, 4, 8, 12, 24, 100, 768, 1024, 2048, 3000, 5000, 1 << 13):
    dom = O.get_evaluation_domain(ms)
    m, kind = zk.domain_size(ms)
    assert (m, kind) == (dom.m, dom.kind), (ms, m, kind)
    v = [rng.randrange(O.R_MOD) for _ in range(m)]
    for op, f in (("FFT", dom.FFT), ("iFFT", dom.iFFT), ("cosetFFT", lambda x: dom.cosetFFT(x, 5)), ("icosetFFT", lambda x: dom.icosetFFT(x, 5)),
                  ("divide_by_Z_on_coset", dom.divide_by_Z_on_coset)):
        got = fl(zk.domain_op(ms, op, fb(v)))
        assert got == f(v), (ms, op)
print("domains small ok", flush=True)
if Rf.available("kernels"):
    for ms in (1 << 16, 196608, 1 << 18, 167275, 1 << 20):
        m, kind = zk.domain_size(ms)
        assert (m, kind) == Rf.domain_size(ms), (ms, m, kind, Rf.domain_size(ms))
        raw = os.urandom(32 * m)
        raw = fb([x % O.R_MOD for x in fl(raw)])
        for op in ("FFT", "iFFT", "cosetFFT", "icosetFFT", "divide_by_Z_on_coset"):
            t = time.time(); got = zk.domain_op(ms, op, raw); t1 = time.time() - t
            exp = Rf.domain_op_bytes(ms, op, raw)
            assert got == exp, (ms, op)
        print("domain", ms, m, kind, "ok vs reference", flush=True)

# 3. MSM vs reference
if Rf.available("kernels"):
    for n in (1, 2, 33, 1000, 1 << 14):
        b1 = Rf.g1_bases_bytes(n, 12345); b2 = Rf.g2_bases_bytes(n, 777)
        sc = [rng.randrange(O.R_MOD) for _ in range(n)]
        for i in range(0, n, 3): sc[i] = rng.choice([0, 1, rng.randrange(1 << 33), O.R_MOD - 1])
        s = fb(sc)
        assert zk.msm_g1(b1, s) == Rf.msm_g1_bytes(b1, s, 1)[0], ("g1", n)
        assert zk.msm_g2(b2, s) == Rf.msm_g2_bytes(b2, s, 1)[0], ("g2", n)
        print("msm", n, "ok", flush=True)

# 4. full prover vs golden
for c in sys.argv[1:] or ["mint"]:
    g = json.load(open(os.path.join(ROOT, "tests/golden/%s.json" % c)))
    w = zlib.decompress(open(os.path.join(ROOT, "tests/golden/%s_assignment.bin.z" % c), "rb").read())
    t = time.time(); pk = zk.ProvingKey(os.path.join(Rf.KEY_DIR, c + "pk.txt")); 
    print(c, "pk load %.2fs (parse %.2f, decompress %.2f)" % (pk.load_seconds, pk.parse_seconds, pk.decompress_seconds), pk.domain_size, pk.domain_kind, pk.num_coefficients, flush=True)
    H, sat = pk.qap_witness_map(w)
    print("  sat", sat, "H sha ok:", hashlib.sha256(H).hexdigest() == g["H_sha256"], flush=True)
    r, s = int(g["r"], 16), int(g["s"], 16)
    res = pk.prove(w, r, s)
    parts = res["parts"]
    for name, lo, hi in (("At", 0, 64), ("Bt_g", 64, 192), ("Bt_h", 192, 256), ("Ht", 256, 320), ("Lt", 320, 384)):
        print("  ", name, parts[lo:hi].hex() == g[name])
    print("  proof equal:", res["proof_hex"] == g["proof_hex"], "timings", res["timings_ms"], "launches", res["launches"], flush=True)
    for _ in range(3):
        t = time.time(); res = pk.prove(w, r, s); dt = time.time() - t
        print("   prove wall %.2f ms" % (dt * 1e3), res["timings_ms"])
    pk.close()
print("imad peak T/s: lo", zk.lib.zkb200_bench_imad_peak(0), "wide", zk.lib.zkb200_bench_imad_peak(1))
