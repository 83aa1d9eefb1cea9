"""Golden proofs of SYNTHETIC transactions from the UNMODIFIED reference prover (oracle/_ref), committed as tests/golden/synthetic.json.

Run in the build container:   python tests/golden/make_synthetic_golden.py
For every circuit and seeds 101..103 (fixtures.synthetic: random secrets, values in [2^32, 2^40), deposit = 256-leaf Merkle tree with
cmtS at a seeded index -- BASELINE.json configs[2]) it records the gen<Circuit>proof arguments, the pinned random_device words and
the 512-character proof the reference prover returns for them, plus the reference verifier's verdict on that proof.
The GPU parity test feeds the same arguments and words through libzkb200's cgo surface and expects the same 512 characters.
"""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import fixtures as F
from oracle import refapi as Rf, bn254_oracle as O

SEEDS = (101, 102, 103)


def verify_args(c, a):
    """verify<Circuit>proof arguments for a gen<Circuit>proof argument list (same mapping as blockmaze_b200.api.verify_args; the deposit
    root comes from the oracle's Merkle restatement)."""
    if c in ("mint", "redeem"):
        return [a[6], a[2], a[7], a[8]]
    if c == "send":
        return [a[5], a[2], a[4], a[11]]
    leaves = [O.uint256S(a[14][66 * i:66 * i + 66]) for i in range(a[15])]
    level = leaves + [bytes(32)] * (256 - len(leaves))
    while len(level) > 1:
        level = [O.sha256_compress(level[i], level[i + 1]) for i in range(0, len(level), 2)]
    return [O.arg_hex(level[0]), a[11], a[8], a[2], a[9], a[6]]


def main():
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    out = {}
    for c in ("mint", "redeem", "send", "deposit"):
        t = time.time(); Rf.load_pk(c, mt=True); t_load = time.time() - t
        rows = []
        for seed in SEEDS:
            args = F.synthetic(c, seed)
            words = O.fixed_rng_words(1000 + seed, 64)
            res = Rf.prove(c, args, words, mt=True)
            assert res["rc"] == 0, (c, seed)
            va = verify_args(c, args)
            ok = Rf.verify(c, res["proof_hex"], va)
            assert ok, (c, seed)
            rows.append(dict(seed=seed, args=args, words=words, proof_hex=res["proof_hex"], verify_args=va, reference_verifies=ok))
        out[c] = rows
        print(c, "ok, pk load %.1fs" % t_load, flush=True)
    json.dump(out, open(os.path.join(HERE, "synthetic.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
