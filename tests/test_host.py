"""CPU-side checks of the product library (no compute on a GPU): the C-ABI loads and exports every symbol the header
declares, the byte-level helpers / witness generators / verifier (all host code) agree with the reference."""
import ctypes as C
import hashlib
import json
import os
import re
import zlib

import pytest

import fixtures as F
from oracle import bn254_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CIRCUITS = ("mint", "send", "deposit", "redeem")


@pytest.fixture(scope="module")
def api():
    import blockmaze_b200.api as api      # loading the library needs no GPU; compute entry points are not called here
    return api


def test_cabi_exports_every_declared_symbol(api):
    header = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(\w+)\s*\(", header)) - {"defined", "if"}
    names = {n for n in names if re.match(r"(zkb200_|gen|verify|compute)", n)}
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(api.lib, n), "libzkb200.so does not export " + n


def test_helpers_known_answers(api):
    z = lambda s: "0x" + s.rjust(64, "0")
    sn_old = api.helper("computePRF", z("1"), z("123456"))
    assert sn_old == "4a31770fe5354a1a9632ebe1481e108cd82ce514ac094c57b5ffdfaea8ac138a"
    sn = api.helper("computePRF", z("1"), z("123"))
    assert sn == "59416ca7b4d0fdcb61dd7fb063db35e9a0a96dd9fecf20a8027aed2d6c4f4006"
    assert api.helper("genCMT", 6, "0x" + sn_old, z("123456")) == "53996012d011396f7a4953c80ee64d5a1b8b7ed676bbea7886b698c09aa991db"
    assert api.helper("genCMT", 13, "0x" + sn, z("123")) == "dbc961ea0d748198f21ea3267534e064eaaea65733dbd362fed93b4748991c27"


def test_helpers_against_oracle_restatement(api):
    import random
    rng = random.Random(4)
    for _ in range(20):
        a, b, c = (bytes(rng.getrandbits(8) for _ in range(32)) for _ in range(3))
        pk = bytes(rng.getrandbits(8) for _ in range(20))
        v = rng.getrandbits(64)
        assert api.helper("computePRF", O.arg_hex(a), O.arg_hex(b)) == O.blob_hex(O.compute_prf(a, b))
        assert api.helper("computeCRH", O.arg_hex(pk), O.arg_hex(b)) == O.blob_hex(O.compute_crh(pk, b))
        assert api.helper("genCMT", v, O.arg_hex(a), O.arg_hex(b)) == O.blob_hex(O.note_cm(v, a, b))
        assert api.helper("genCMTS", v, O.arg_hex(pk), O.arg_hex(b), O.arg_hex(c)) == O.blob_hex(O.notes_cm(v, pk, b, c))


def test_hex_parsing_quirks(api):
    """uint256S semantics (uint256.h:200-226): optional 0x, short strings are right-aligned, parsing stops at the first non-hex."""
    one = api.helper("computePRF", "1", "0x123456")
    assert one == api.helper("computePRF", "0x" + "1".rjust(64, "0"), "0x" + "123456".rjust(64, "0"))
    assert api.helper("computePRF", "  0X1", "123456zz") == one


def test_gen_root_matches_full_tree(api):
    import random
    rng = random.Random(8)
    for n in (0, 1, 2, 3, 16, 255, 256):
        leaves = [bytes(rng.getrandbits(8) for _ in range(32)) for _ in range(n)]
        level = leaves + [bytes(32)] * (256 - n)
        while len(level) > 1:
            level = [O.sha256_compress(level[i], level[i + 1]) for i in range(0, len(level), 2)]
        assert api.helper("genRoot", "".join(O.arg_hex(x) for x in leaves), n) == O.blob_hex(level[0])


@pytest.mark.parametrize("circuit", CIRCUITS)
def test_witness_equals_reference_golden(api, circuit):
    """Native witness generator == assignment produced by the reference gadgets for the reference's own fixture."""
    g = json.load(open(os.path.join(GOLD, circuit + ".json")))
    w = api.witness(circuit, g["args"])
    assert hashlib.sha256(w).hexdigest() == g["assignment_sha256"]
    assert w == zlib.decompress(open(os.path.join(GOLD, circuit + "_assignment.bin.z"), "rb").read())


@pytest.mark.parametrize("circuit", CIRCUITS)
def test_witness_equals_reference_on_synthetic_transactions(api, ref, circuit):
    if not ref.available(circuit):
        pytest.skip("reference circuit harness not built")
    for seed in range(4):
        args = F.synthetic(circuit, seed)
        theirs, sat = ref.witness(circuit, args)
        assert sat
        assert api.witness(circuit, args) == theirs


def test_deposit_witness_leaf_positions(api, ref):
    """Merkle path handling: cmtS at the first, a middle and the last of 256 leaves, and with fewer leaves than the tree holds."""
    if not ref.available("deposit"):
        pytest.skip("reference circuit harness not built")
    import random
    for n, idx in ((256, 0), (256, 255), (256, 100), (5, 4), (1, 0)):
        rng = random.Random(n * 1000 + idx)
        leaves = [bytes(rng.getrandbits(8) for _ in range(32)) for _ in range(n)]
        args = F.deposit_fixture(leaves=leaves, index=idx)
        theirs, sat = ref.witness("deposit", args)
        assert sat and api.witness("deposit", args) == theirs


@pytest.mark.parametrize("circuit", CIRCUITS)
def test_verifier_accepts_reference_proofs_and_rejects_tampering(api, ref, circuit):
    """verify*proof needs only the vk file (1.3 KB); golden proofs come from the reference prover."""
    vk = os.path.join(ref.KEY_DIR, circuit + "vk.txt")
    if not os.path.exists(vk):
        pytest.skip("reference keys not present")
    api.set_key_dir(ref.KEY_DIR)
    g = json.load(open(os.path.join(GOLD, circuit + ".json")))
    va = api.verify_args(circuit, g["args"])
    assert api.verify_proof(circuit, g["proof_hex"], va)
    h = g["proof_hex"]
    assert not api.verify_proof(circuit, h[:200] + ("1" if h[200] != "1" else "2") + h[201:], va)      # not on the curve any more
    assert not api.verify_proof(circuit, h[:384] + h[:128], va)                                           # C := A, a valid point
    bad = list(va)
    bad[1] = bad[1][:-1] + ("1" if bad[1][-1] != "1" else "2")
    assert not api.verify_proof(circuit, h, bad)
    assert not api.verify_proof(circuit, "0" * 63 + "1" + "0" * 63 + "2" + h[128:], va)                  # default-proof style A
    assert not api.verify_proof(circuit, "zz" + h[2:], va)


def test_wallet_generator_equals_fixture_generator(api):
    """bench.py builds its transactions with blockmaze_b200.wallet (the library's own cgo helpers); the reference arm and the parity tests
    use tests/golden/fixtures.py (oracle hashing).  Same seeds must give the same gen*proof argument lists."""
    from blockmaze_b200 import wallet as W
    for c in ("mint", "send", "deposit", "redeem"):
        for seed in (0, 1, 5, 1023):
            assert W.synthetic(c, seed) == F.synthetic(c, seed)
    assert W.mint(13, 6, 7, "1", "123456", "123") == F.mint_fixture()
    assert W.send(14, 22, 8, "1", "123456", "12", "456", "123") == F.send_fixture()


def test_g1_sum_on_host(api):
    """zkb200_g1_sum (the host step of an MSM split by point range) against the oracle's group law, incl. infinity and P + (-P)."""
    import random
    rng = random.Random(3)
    pts = [O.G1.to_affine(O.G1.mul(rng.randrange(1, O.R_MOD), O.G1.from_affine(O.G1_GEN))) for _ in range(5)]
    enc = lambda p: p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little")
    acc = O.G1.zero()
    for p in pts:
        acc = O.G1.add(acc, O.G1.from_affine(p))
    assert api.g1_sum([enc(p) for p in pts] + [bytes(64)]) == enc(O.G1.to_affine(acc))
    neg = (pts[0][0], O.Q_MOD - pts[0][1])
    assert api.g1_sum([enc(pts[0]), enc(neg)]) == bytes(64)
    assert api.g1_sum([enc(pts[1]), enc(pts[1])]) == enc(O.G1.to_affine(O.G1.add(O.G1.from_affine(pts[1]), O.G1.from_affine(pts[1]))))


def test_equal_range_accumulation_index_math():
    """Host restatement of the index arithmetic of msm_accumulate_kernel / msm_fold_* (csrc/msm.cuh): the bucket-sorted entry list is cut
    into equal ranges, thread t stores the piece of bucket b in slot t + b, bucket b later folds slots off[b]/L + b .. (off[b+1]-1)/L + b.
    Checks that slots never collide, that the fold sees exactly the pieces of its bucket, and that every entry is counted once."""
    import random
    rng = random.Random(11)
    for trial in range(30):
        nb = rng.choice([5, 64, 300])
        counts = [rng.choice([0, 0, 1, 2, 7, 40, rng.randrange(0, 3000)]) for _ in range(nb)]
        off = [0]
        for c in counts:
            off.append(off[-1] + c)
        total, threads = off[-1], rng.choice([128, 256, 1024])
        L = max(16, -(-total // threads))
        slots = {}
        for t in range(threads):
            e0, e1 = t * L, min(t * L + L, total)
            if e0 >= total:
                continue
            b = max(i for i in range(nb) if off[i] <= e0)           # the kernel's binary search: last b with off[b] <= e0
            assert off[b + 1] > e0
            acc, nxt = 0, off[b + 1]
            for e in range(e0, e1):
                if e == nxt:
                    assert (t + b) not in slots
                    slots[t + b] = (b, acc)
                    acc = 0
                    b = max(i for i in range(nb) if off[i] <= e)
                    nxt = off[b + 1]
                acc += 1
            assert (t + b) not in slots
            slots[t + b] = (b, acc)
        assert max(slots, default=0) <= threads + nb
        for b in range(nb):
            if counts[b] == 0:
                continue
            t0, t1 = off[b] // L, (off[b + 1] - 1) // L
            pieces = [slots.pop(t + b) for t in range(t0, t1 + 1)]
            assert all(pb == b for pb, _ in pieces) and sum(c for _, c in pieces) == counts[b]
        assert not slots


def test_verifier_concurrent_calls(api, ref):
    """verify*proof from several threads at once (the key is shared without a lock): every reference proof verifies, a proof whose C was
    replaced by its A (a valid point) fails."""
    from concurrent.futures import ThreadPoolExecutor
    if not os.path.exists(os.path.join(ref.KEY_DIR, "mintvk.txt")):
        pytest.skip("reference keys not present")
    api.set_key_dir(ref.KEY_DIR)
    jobs = []
    for c in ("mint", "send", "deposit", "redeem"):
        g = json.load(open(os.path.join(ROOT, "tests", "golden", c + ".json")))
        va = api.verify_args(c, g["args"])
        bad = g["proof_hex"][:384] + g["proof_hex"][:128]
        jobs += [(c, g["proof_hex"], va, True), (c, bad, va, False)] * 3
    with ThreadPoolExecutor(6) as pool:
        got = list(pool.map(lambda j: api.verify_proof(j[0], j[1], j[2]), jobs))
    assert got == [j[3] for j in jobs]


@pytest.mark.parametrize("circuit", CIRCUITS)
def test_verifier_agrees_with_reference_verifier_incl_noncanonical_encodings(api, ref, circuit):
    """Same verdict as the reference verifier (verify_<c>_proof -> r1cs_gg_ppzksnark_verifier_strong_IC, run on an explicit vk path by
    oracle/ref_harness) on: the golden proof; tampered proofs; and NON-CANONICAL coordinate encodings x + k*q < 2^256.  The reference
    builds Fq from the raw 256-bit integer (Fp_model(bigint), fp.tcc:190-194), which reduces mod q, so it ACCEPTS those -- a drop-in
    verifier must too, or a mixed network forks on such a transaction."""
    if not ref.available(circuit) or not os.path.exists(os.path.join(ref.KEY_DIR, circuit + "vk.txt")):
        pytest.skip("reference circuit harness / keys not present")
    api.set_key_dir(ref.KEY_DIR)
    g = json.load(open(os.path.join(GOLD, circuit + ".json")))
    va = api.verify_args(circuit, g["args"])
    h = g["proof_hex"]
    q = O.Q_MOD

    def bump(proof, coord, k):
        v = int(proof[64 * coord:64 * coord + 64], 16) + k * q
        assert v < 1 << 256
        return proof[:64 * coord] + "%064x" % v + proof[64 * coord + 64:]

    cases = [h, bump(h, 0, 1), bump(h, 1, 1), bump(h, 2, 1), bump(h, 5, 2), bump(h, 6, 3), bump(bump(h, 7, 4), 0, 4),
             h[:384] + h[:128],                                       # C := A
             bump(h[:384] + h[:128], 6, 1),
             h[:200] + ("1" if h[200] != "1" else "2") + h[201:],     # off the curve
             "f" * 64 + h[64:]]                                       # 2^256 - 1 as A.x
    bad = list(va)
    bad[0] = bad[0][:-1] + ("1" if bad[0][-1] != "1" else "2")
    verdicts = ref.verify_many([(circuit, p, va, None) for p in cases] + [(circuit, h, bad, None)])
    assert verdicts[:7] == [True] * 7 and not any(verdicts[7:])       # pins the reference's behaviour
    assert [api.verify_proof(circuit, p, va) for p in cases] + [api.verify_proof(circuit, h, bad)] == verdicts


def test_device_scheduler_policy(api):
    """zkb::DeviceSched (csrc/sched.hpp), the policy gen*proof uses to spread proofs over the GPUs of one box: least proofs in flight,
    ties round-robin.  Host-only, so it is checked here without a GPU."""
    import random
    lib = api.lib
    s = lib.zkb200_sched_create(4)
    try:
        first = [lib.zkb200_sched_pick(s) for _ in range(8)]
        assert first == [0, 1, 2, 3, 0, 1, 2, 3]                      # an idle box fills in rotation
        assert [lib.zkb200_sched_inflight(s, d) for d in range(4)] == [2, 2, 2, 2]
        lib.zkb200_sched_done(s, 2); lib.zkb200_sched_done(s, 2)
        assert lib.zkb200_sched_pick(s) == 2 and lib.zkb200_sched_pick(s) == 2      # the drained device takes the next two
        held = [0, 1, 2, 3, 0, 1, 2, 3]
        rng = random.Random(5)
        for _ in range(2000):                                          # random arrivals and completions: load never differs by more than one ...
            if held and rng.random() < 0.5:
                d = held.pop(rng.randrange(len(held)))
                lib.zkb200_sched_done(s, d)
            else:
                load = [lib.zkb200_sched_inflight(s, d) for d in range(4)]
                d = lib.zkb200_sched_pick(s)
                assert load[d] == min(load)                            # ... because a pick always goes to a least-loaded device
                held.append(d)
        for d in held:
            lib.zkb200_sched_done(s, d)
        assert [lib.zkb200_sched_inflight(s, d) for d in range(4)] == [0, 0, 0, 0]
        lib.zkb200_sched_done(s, 1)                                    # spurious completion is ignored
        assert lib.zkb200_sched_inflight(s, 1) == 0 and lib.zkb200_sched_inflight(s, 9) == -1
    finally:
        lib.zkb200_sched_free(s)
    # concurrent callers (goroutines in geth): totals add up and stay balanced
    from concurrent.futures import ThreadPoolExecutor
    s = lib.zkb200_sched_create(8)
    counts = [0] * 8

    def job(_):
        d = lib.zkb200_sched_pick(s)
        counts[d] += 1
        lib.zkb200_sched_done(s, d)
    with ThreadPoolExecutor(16) as pool:
        list(pool.map(job, range(4000)))
    assert sum(counts) == 4000 and all(lib.zkb200_sched_inflight(s, d) == 0 for d in range(8))
    lib.zkb200_sched_free(s)


def test_prove_batch_rejects_bad_arguments_without_touching_a_gpu(api):
    assert api.lib.zkb200_prove_batch(0, None, None, 0) == 0
    t = (api.Tx * 1)()
    t[0].circuit = 7
    out = C.create_string_buffer(513)
    assert api.lib.zkb200_prove_batch(1, t, out, 1) == -1
    assert api.lib.zkb200_prove_batch(1, None, out, 1) == -1
    tx = api._tx("deposit", F.deposit_fixture())
    assert tx.circuit == 2 and tx.n == 16 and list(tx.u) == [264, 255, 9] and tx.s[11].startswith(b"0x") and len(tx.s[11]) == 16 * 66
    tx = api._tx("send", F.send_fixture())
    assert tx.circuit == 1 and list(tx.u) == [22, 8, 14] and tx.s[10] is not None and tx.s[11] is None


def test_random_words_hook_needs_test_env(api):
    """zkb200_set_random_words pins (r, s): a production process (no ZKB200_TEST_RNG=1) cannot switch it on."""
    import subprocess, sys
    code = ("import os, sys; os.environ.pop('ZKB200_TEST_RNG', None); sys.path.insert(0, %r)\n"
            "import ctypes as C; from blockmaze_b200 import api\n"
            "w = (C.c_uint32 * 16)(*range(16))\n"
            "print('RC', api.lib.zkb200_set_random_words(C.cast(w, C.c_void_p), 16), api.lib.zkb200_set_random_words(None, 0))\n" % ROOT)
    env = {k: v for k, v in os.environ.items() if k != "ZKB200_TEST_RNG"}
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert "RC -1 0" in out.stdout, out.stdout + out.stderr
    w = (C.c_uint32 * 16)(*range(16))
    assert api.lib.zkb200_set_random_words(C.cast(w, C.c_void_p), 16) == 0          # this process has the variable (tests/conftest.py)
    assert api.lib.zkb200_set_random_words(None, 0) == 0


def _batch_items():
    """Reference proofs of all four circuits: the reference fixtures plus the synthetic goldens (16 proofs, mixed circuits)."""
    items = []
    syn = json.load(open(os.path.join(GOLD, "synthetic.json")))
    for c in CIRCUITS:
        g = json.load(open(os.path.join(GOLD, c + ".json")))
        items.append((c, g["proof_hex"], g["args"]))
        items += [(c, row["proof_hex"], row["args"]) for row in syn[c]]
    return items


def test_batch_verifier_matches_per_proof_and_reference_verdicts(api, ref):
    """zkb200_verify_batch (random linear combination, one final exponentiation per batch) gives the verdicts of
    verify<Circuit>proof and of the reference verifier: an all-valid mixed batch, batches with one / several bad proofs of every kind
    (wrong inputs, a valid point swapped in, off-curve, malformed hex, non-canonical but valid encodings), single-item and empty batches."""
    if not os.path.exists(os.path.join(ref.KEY_DIR, "mintvk.txt")) or not ref.available("mint"):
        pytest.skip("reference keys / harness not present")
    api.set_key_dir(ref.KEY_DIR)
    good = [(c, p, api.verify_args(c, a)) for c, p, a in _batch_items()]
    assert api.verify_batch(good) == [True] * len(good)
    assert api.verify_batch(good, threads=1) == [True] * len(good)
    assert api.verify_batch([]) == [] and api.verify_batch(good[:1]) == [True]
    q = O.Q_MOD
    mixed = list(good)
    c, p, va = mixed[2]
    mixed[2] = (c, p[:384] + p[:128], va)                                                       # C := A (valid point, wrong proof)
    c, p, va = mixed[5]
    mixed[5] = (c, p, [va[0][:-1] + ("1" if va[0][-1] != "1" else "2")] + list(va[1:]))       # wrong public input
    c, p, va = mixed[7]
    mixed[7] = (c, "%064x" % (int(p[:64], 16) + q) + p[64:], va)                                # A.x + q: valid for the reference, so valid here
    c, p, va = mixed[9]
    mixed[9] = (c, p[:200] + ("1" if p[200] != "1" else "2") + p[201:], va)                    # B off the curve
    c, p, va = mixed[11]
    mixed[11] = (c, "zz" + p[2:], va)                                                           # not hex
    c, p, va = mixed[12]
    mixed[12] = (mixed[13][0], p, mixed[13][2]) if mixed[13][0] == c else (c, mixed[13][1], va)   # another transaction's proof
    want = [api.verify_proof(c, p, va) for c, p, va in mixed]
    assert want.count(False) == 5 and want[7] is True
    assert api.verify_batch(mixed) == want
    assert ref.verify_many([(c, p, va, None) for c, p, va in mixed]) == want
    # a B that is on the twist curve but outside the order-r subgroup goes to the per-proof path and gets its verdict
    c, p, va = good[0]
    x = 1
    while True:                                           # a point of E'(Fq2) from a small x; its order divides #E' = r * (2q - r), almost surely not r
        x += 1
        try:
            y = O.fq2_sqrt(O.Fq2(x, 0) * O.Fq2(x, 0) * O.Fq2(x, 0) + O.TWIST_B)
            break
        except ValueError:
            continue
    assert O.G2.mul(O.R_MOD, O.G2.from_affine((O.Fq2(x, 0), y))) != O.G2.zero()          # really outside the subgroup
    hx = lambda v: "%064x" % v
    off = p[:128] + hx(0) + hx(x) + hx(y.c1) + hx(y.c0) + p[384:]
    assert api.verify_batch(good[:3] + [(c, off, va)] + good[3:6]) == [True] * 3 + [api.verify_proof(c, off, va)] + [True] * 3


def test_batch_verifier_is_faster_than_one_by_one(api, ref):
    import time
    if not os.path.exists(os.path.join(ref.KEY_DIR, "mintvk.txt")):
        pytest.skip("reference keys not present")
    api.set_key_dir(ref.KEY_DIR)
    good = [(c, p, api.verify_args(c, a)) for c, p, a in _batch_items()] * 4            # 64 proofs
    t0 = time.perf_counter(); assert all(api.verify_proof(c, p, va) for c, p, va in good); t1 = time.perf_counter()
    assert all(api.verify_batch(good, threads=1)); t2 = time.perf_counter()
    print("one by one %.1f ms/proof, batch (1 thread) %.1f ms/proof" % (1e3 * (t1 - t0) / len(good), 1e3 * (t2 - t1) / len(good)))
    assert (t2 - t1) < 0.8 * (t1 - t0)


def test_g2_subgroup_fast_test_equals_definition(api):
    """The batch verifier's membership test psi(Q) == [6z^2]Q against the definition [r]Q == O: multiples of the generator are in, points
    of the twist curve built from arbitrary x (cofactor 2q - r, so almost never in G2) are out, and the two tests never disagree."""
    import random
    rng = random.Random(9)
    enc = lambda P: b"".join(int(v).to_bytes(32, "little") for v in (P[0].c0, P[0].c1, P[1].c0, P[1].c1))
    gen = (O.Fq2(*O.G2_GEN[0]), O.Fq2(*O.G2_GEN[1]))
    for k in (1, 2, 12345, rng.randrange(O.R_MOD), O.R_MOD - 1):
        P = O.G2.to_affine(O.G2.mul(k, O.G2.from_affine(gen)))
        assert api.lib.zkb200_g2_subgroup_check(enc(P)) == 3
    outside = 0
    while outside < 12:
        x = O.Fq2(rng.randrange(O.Q_MOD), rng.randrange(O.Q_MOD))
        try:
            y = O.fq2_sqrt(x * x * x + O.TWIST_B)
        except ValueError:
            continue
        got = api.lib.zkb200_g2_subgroup_check(enc((x, y)))
        assert got in (0, 3)
        outside += got == 0
        # clearing the cofactor lands in G2 again
        if outside == 1:
            h = 2 * O.Q_MOD - O.R_MOD
            assert api.lib.zkb200_g2_subgroup_check(enc(O.G2.to_affine(O.G2.mul(h, O.G2.from_affine((x, y)))))) == 3


@pytest.mark.parametrize("circuit", CIRCUITS)
def test_deferred_sha256_runs_expand_to_the_same_assignment(api, circuit):
    """GPU witness path, host half: the generator in deferred mode (SHA-256 compression runs left out, 100-byte seeds recorded) plus the
    seed expansion of csrc/witness_sha.hpp -- the very code the GPU kernel runs -- reproduces the full assignment variable for variable,
    for the reference fixture (golden sha256) and for synthetic transactions (deposit: 18 compressions, 256-leaf tree)."""
    g = json.load(open(os.path.join(GOLD, circuit + ".json")))
    w = api.witness(circuit, g["args"], defer=True)
    assert hashlib.sha256(w).hexdigest() == g["assignment_sha256"]
    for seed in (3, 4):
        args = F.synthetic(circuit, seed)
        assert api.witness(circuit, args, defer=True) == api.witness(circuit, args)
