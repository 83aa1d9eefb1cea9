"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile).

Run in the build container (needs oracle/_ref incl. prfKey):   python tests/golden/make_golden.py [circuit ...]
Per circuit it writes
  tests/golden/<c>.json              args, pinned random words, r, s, sha256 of assignment / coefficients_for_H,
                                     the five MSM results, the 512-char proof, reference phase timings of this run
  tests/golden/<c>_assignment.bin.z  zlib of the full variable assignment (n x 32 B LE canonical) from the reference gadgets
  oracle/_ref/dumps/<c>/*            full binary dumps (git-ignored; travel to the GPU box)
With --cgo it additionally calls the reference's real gen<C>proof() from libzk_<c>.so in a child process under
LD_PRELOAD=libfixed_rng.so and records that the harness proof equals it byte for byte (pins the harness itself).
"""
import hashlib, json, os, subprocess, sys, time, zlib
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
import fixtures as F
from oracle import refapi as Rf, bn254_oracle as O

SEED = 42

CGO_CHILD = r'''
import ctypes as C, sys, json
sys.path.insert(0, %(root)r)
from oracle import refapi as Rf
c, args = %(circuit)r, json.loads(%(args)r)
L = C.CDLL(%(lib)r)
f = getattr(L, "gen%%sproof" %% c.capitalize()); f.restype = C.c_void_p; f.argtypes = Rf.SIGS[c]
p = f(*Rf._enc(args))
print("PROOF", C.string_at(p, 512).decode())
'''


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    circuits = argv or ["mint", "redeem", "send", "deposit"]
    for c in circuits:
        args = F.FIXTURES[c]()
        words = O.fixed_rng_words(SEED, 64)
        r, s = O.fr_from_words(words)
        outdir = os.path.join(Rf.REF_DIR, "dumps", c)
        os.makedirs(outdir, exist_ok=True)
        t = time.time(); Rf.load_pk(c); t_load = time.time() - t
        res = Rf.prove(c, args, words, outdir)
        assert res["rc"] == 0, res
        assign = open(os.path.join(outdir, "assignment.bin"), "rb").read()
        Hc = open(os.path.join(outdir, "coefficients_for_H.bin"), "rb").read()
        rs = open(os.path.join(outdir, "rs.bin"), "rb").read()
        assert Rf.fr_list(rs) == [r, s], "oracle fr_from_words disagrees with the reference harness"
        wit, sat = Rf.witness(c, args)
        assert sat and wit == assign
        g = dict(circuit=c, args=args, seed=SEED, words=words, r="%064x" % r, s="%064x" % s,
                 num_variables=len(assign) // 32, domain_size=len(Hc) // 32 - 1,
                 assignment_sha256=hashlib.sha256(assign).hexdigest(), H_sha256=hashlib.sha256(Hc).hexdigest(),
                 At=res["pts"][0:64].hex(), Bt_g=res["pts"][64:192].hex(), Bt_h=res["pts"][192:256].hex(),
                 Ht=res["pts"][256:320].hex(), Lt=res["pts"][320:384].hex(), proof_hex=res["proof_hex"],
                 reference_timings_1thread=dict(zip(["qap", "A", "B", "H", "L"], res["timings"])), reference_pk_load_s=t_load)
        if "--cgo" in sys.argv:
            env = dict(os.environ, LD_PRELOAD=os.path.join(Rf.REF_DIR, "libfixed_rng.so"), ZK_FIXED_SEED=str(SEED))
            code = CGO_CHILD % dict(root=ROOT, circuit=c, args=json.dumps(args), lib=os.path.join(Rf.REF_DIR, "libzk_%s.so" % c))
            out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout
            cgo = [l for l in out.splitlines() if l.startswith("PROOF ")][0].split()[1]
            g["cgo_genproof_equal"] = (cgo == res["proof_hex"])
            assert g["cgo_genproof_equal"], (cgo, res["proof_hex"])
        json.dump(g, open(os.path.join(HERE, c + ".json"), "w"), indent=1)
        open(os.path.join(HERE, c + "_assignment.bin.z"), "wb").write(zlib.compress(assign, 9))
        print(c, "ok", g["reference_timings_1thread"], "pk load %.1fs" % t_load, flush=True)


if __name__ == "__main__":
    main()
