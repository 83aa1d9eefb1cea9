"""TEST INFRASTRUCTURE -- ctypes access to the UNMODIFIED reference built into oracle/_ref (see oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Wire formats: field elements 32-byte LE canonical; G1 affine 64 B; G2 affine 128 B (x.c0 x.c1 y.c0 y.c1); zero bytes = infinity.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
KEY_DIR = os.path.join(REF_DIR, "prfKey")

_libs = {}


def available(name="kernels"):
    return os.path.exists(os.path.join(REF_DIR, "libref_%s.so" % name))


def lib(name):
    """name: kernels | kernels_mt | mint | send | deposit | redeem (| *_mt)."""
    if name not in _libs:
        path = os.path.join(REF_DIR, "libref_%s.so" % name)
        if not os.path.exists(path):
            raise FileNotFoundError("%s missing: run `make -C oracle` where /root/reference exists" % path)
        _libs[name] = C.CDLL(path, mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
    return _libs[name]


def fr_bytes(vals): return b"".join(int(v).to_bytes(32, "little") for v in vals)
def fr_list(buf): return [int.from_bytes(buf[i:i + 32], "little") for i in range(0, len(buf), 32)]


def g1_bytes(a):
    return bytes(64) if a is None else int(a[0]).to_bytes(32, "little") + int(a[1]).to_bytes(32, "little")


def g1_from(buf):
    if not any(buf):
        return None
    return (int.from_bytes(buf[:32], "little"), int.from_bytes(buf[32:64], "little"))


def g2_bytes(a):
    if a is None:
        return bytes(128)
    x, y = a
    return b"".join(int(v).to_bytes(32, "little") for v in (x.c0, x.c1, y.c0, y.c1))


def g2_from(buf):
    from . import bn254_oracle as O
    if not any(buf):
        return None
    v = [int.from_bytes(buf[i:i + 32], "little") for i in range(0, 128, 32)]
    return (O.Fq2(v[0], v[1]), O.Fq2(v[2], v[3]))


# ---- kernels ---------------------------------------------------------------------------------------------------
def domain_size(min_size, mt=False):
    k = C.c_int(0)
    L = lib("kernels_mt" if mt else "kernels")
    L.ref_domain_size.restype = C.c_long
    m = L.ref_domain_size(C.c_size_t(min_size), C.byref(k))
    return m, {0: "basic_radix2", 1: "step_radix2"}.get(k.value, "other")


OPS = {"FFT": 0, "iFFT": 1, "cosetFFT": 2, "icosetFFT": 3, "divide_by_Z_on_coset": 4}


def domain_op_bytes(min_size, op, data, mt=False):
    L = lib("kernels_mt" if mt else "kernels")
    L.ref_domain_op.restype = C.c_long
    buf = C.create_string_buffer(bytes(data), len(data))
    m = L.ref_domain_op(C.c_size_t(min_size), OPS[op], buf, C.c_size_t(len(data) // 32))
    if m < 0:
        raise ValueError("ref_domain_op failed")
    return buf.raw


def domain_op_timed(min_size, op, data, mt=False):
    """(transformed bytes, seconds of the transform alone)."""
    L = lib("kernels_mt" if mt else "kernels")
    L.ref_domain_op_timed.restype = C.c_long
    buf = C.create_string_buffer(bytes(data), len(data))
    sec = C.c_double(0)
    m = L.ref_domain_op_timed(C.c_size_t(min_size), OPS[op], buf, C.c_size_t(len(data) // 32), C.byref(sec))
    if m < 0:
        raise ValueError("ref_domain_op failed")
    return buf.raw, sec.value


def threads(mt=True):
    return int(lib("kernels_mt" if mt else "kernels").ref_threads())


def domain_op(min_size, op, vals, mt=False):
    return fr_list(domain_op_bytes(min_size, op, fr_bytes(vals), mt))


def domain_element(min_size, idx):
    L = lib("kernels")
    L.ref_domain_element.restype = C.c_long
    out = C.create_string_buffer(32)
    L.ref_domain_element(C.c_size_t(min_size), C.c_size_t(idx), out)
    return int.from_bytes(out.raw, "little")


def msm_g1_bytes(bases, scalars, method=0, chunks=1, mt=False):
    L = lib("kernels_mt" if mt else "kernels")
    n = len(scalars) // 32
    out = C.create_string_buffer(64)
    sec = C.c_double(0)
    L.ref_msm_g1(C.c_size_t(n), bytes(bases), bytes(scalars), method, C.c_size_t(chunks), out, C.byref(sec))
    return out.raw, sec.value


def msm_g2_bytes(bases, scalars, method=0, chunks=1, mt=False):
    L = lib("kernels_mt" if mt else "kernels")
    n = len(scalars) // 32
    out = C.create_string_buffer(128)
    sec = C.c_double(0)
    L.ref_msm_g2(C.c_size_t(n), bytes(bases), bytes(scalars), method, C.c_size_t(chunks), out, C.byref(sec))
    return out.raw, sec.value


def g1_bases_bytes(n, step=7):
    L = lib("kernels")
    out = C.create_string_buffer(64 * n)
    L.ref_g1_bases(C.c_size_t(n), int(step).to_bytes(32, "little"), out)
    return out.raw


def g2_bases_bytes(n, step=7):
    L = lib("kernels")
    out = C.create_string_buffer(128 * n)
    L.ref_g2_bases(C.c_size_t(n), int(step).to_bytes(32, "little"), out)
    return out.raw


def sha512_rng(first, n):
    """libff::SHA512_rng<Fr>(first + i), i < n, as canonical ints."""
    out = C.create_string_buffer(32 * n)
    lib("kernels").ref_sha512_rng(C.c_uint64(first), C.c_size_t(n), out)
    return fr_list(out.raw)


def g1_mul(a, k):
    out = C.create_string_buffer(64)
    lib("kernels").ref_g1_mul(g1_bytes(a), int(k).to_bytes(32, "little"), out)
    return g1_from(out.raw)


def g2_mul(a, k):
    out = C.create_string_buffer(128)
    lib("kernels").ref_g2_mul(g2_bytes(a), int(k).to_bytes(32, "little"), out)
    return g2_from(out.raw)


def field_op(which, op, a, b=None):
    """which: 'fr'|'fq'; op: mul|add|sub|inv on lists of canonical ints."""
    n = len(a)
    out = C.create_string_buffer(32 * n)
    lib("kernels").ref_field_op(0 if which == "fr" else 1, {"mul": 0, "add": 1, "sub": 2, "inv": 3}[op], C.c_size_t(n),
                                fr_bytes(a), fr_bytes(b) if b is not None else None, out)
    return fr_list(out.raw)


def to_mont(which, a):
    out = C.create_string_buffer(32 * len(a))
    lib("kernels").ref_to_mont(0 if which == "fr" else 1, C.c_size_t(len(a)), fr_bytes(a), out)
    return fr_list(out.raw)


def qap_witness_map(cs, full, mt=False):
    """cs: bn254_oracle.R1CS.  Returns (coefficients_for_H list of m+1, seconds)."""
    import numpy as np
    L = lib("kernels_mt" if mt else "kernels")
    L.ref_qap_witness_map.restype = C.c_long
    args = []
    for w in range(3):
        rowptr, col, coef = [0], [], []
        for row in cs.rows:
            for idx, c in row[w]:
                col.append(idx)
                coef.append(c)
            rowptr.append(len(col))
        args += [np.asarray(rowptr, dtype=np.uint32).tobytes(), np.asarray(col, dtype=np.uint32).tobytes(), fr_bytes(coef)]
    nc = cs.num_constraints
    cap = (4 * (nc + cs.num_inputs + 1) + 8) * 32
    out = C.create_string_buffer(cap)
    sec = C.c_double(0)
    m = L.ref_qap_witness_map(C.c_size_t(cs.num_inputs), C.c_size_t(cs.num_variables), C.c_size_t(nc), *args,
                              fr_bytes(full), out, C.c_size_t(cap), C.byref(sec))
    if m < 0:
        raise ValueError("ref_qap_witness_map failed")
    return fr_list(out.raw[:(m + 1) * 32]), sec.value


# ---- circuits --------------------------------------------------------------------------------------------------
SIGS = {
    # ctypes argument types of gen<Circuit>proof (SRC/<c>/<c>cgo.hpp)
    "mint": [C.c_uint64, C.c_uint64] + [C.c_char_p] * 6 + [C.c_uint64, C.c_char_p],
    "redeem": [C.c_uint64, C.c_uint64] + [C.c_char_p] * 6 + [C.c_uint64, C.c_char_p],
    "send": [C.c_uint64] + [C.c_char_p] * 5 + [C.c_uint64, C.c_char_p, C.c_uint64] + [C.c_char_p] * 5,
    "deposit": [C.c_uint64, C.c_uint64] + [C.c_char_p] * 8 + [C.c_uint64] + [C.c_char_p] * 4 + [C.c_int, C.c_char_p, C.c_char_p],
}
NUM_VARS = {"mint": 151512, "send": 227046, "deposit": 457127, "redeem": 151579}


def _enc(args):
    return [a.encode() if isinstance(a, str) else a for a in args]


def witness(circuit, args):
    """Full variable assignment (primary || aux) computed by the reference gadget code.  Returns (bytes, satisfied)."""
    L = lib(circuit)
    f = getattr(L, "ref_%s_witness" % circuit)
    f.restype = C.c_long
    f.argtypes = SIGS[circuit] + [C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
    n = NUM_VARS[circuit]
    out = C.create_string_buffer(32 * n)
    sat = C.c_int(0)
    got = f(*_enc(args), C.cast(out, C.c_void_p), n, C.byref(sat))
    if sat.value and got != n:
        raise ValueError("unexpected variable count %d" % got)
    return (out.raw if sat.value else b""), bool(sat.value)


def load_pk(circuit, path=None, mt=False):
    L = lib(circuit + ("_mt" if mt else ""))
    path = path or os.path.join(KEY_DIR, circuit + "pk.txt")
    return getattr(L, "ref_%s_load_pk" % circuit)(path.encode())


def prove(circuit, args, words, outdir=None, mt=False):
    """Reference prover with pinned randomness.  Returns dict(proof_hex, pts(bytes 384), timings[5], rc)."""
    L = lib(circuit + ("_mt" if mt else ""))
    f = getattr(L, "ref_%s_prove" % circuit)
    f.argtypes = SIGS[circuit] + [C.c_void_p, C.c_size_t, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
    W = (C.c_uint32 * len(words))(*words)
    hexbuf = C.create_string_buffer(513)
    pts = C.create_string_buffer(384)
    tim = (C.c_double * 5)()
    rc = f(*_enc(args), C.cast(W, C.c_void_p), len(words), outdir.encode() if outdir else None,
           C.cast(hexbuf, C.c_void_p), C.cast(pts, C.c_void_p), C.cast(tim, C.c_void_p))
    return dict(rc=rc, proof_hex=hexbuf.value.decode(), pts=pts.raw, timings=list(tim))


VERIFY_SIGS = {"mint": [C.c_char_p] * 5 + [C.c_uint64], "redeem": [C.c_char_p] * 5 + [C.c_uint64], "send": [C.c_char_p] * 6, "deposit": [C.c_char_p] * 8}


def _verify_here(circuit, proof_hex, verify_args, vk_path):
    L = lib(circuit)
    f = getattr(L, "ref_%s_verify" % circuit)
    f.restype = C.c_int
    f.argtypes = VERIFY_SIGS[circuit]
    return bool(f(vk_path.encode(), proof_hex.encode(), *_enc(verify_args)))


# The reference's key parsers read decimal text with std::istream, and oracle/_ref/libff.so carries its own instantiations of the
# libstdc++ locale facets (GNU-unique symbols): they only work when libff.so enters the process BEFORE any other copy of libstdc++
# (libzkb200.so, numpy, torch all bring one).  Anything that parses a key file -- verification, the live prover -- therefore runs in
# a fresh interpreter that loads nothing but the reference.
_CHILD = ("import sys, json\nsys.path.insert(0, %r)\nfrom oracle import refapi as R\n"
          "print('RESULT ' + json.dumps(R._child(json.loads(sys.stdin.read()))))\n")


def _child(req):
    if req["op"] == "verify":
        return [_verify_here(c, p, a, vk) for c, p, a, vk in req["jobs"]]
    if req["op"] == "prove":
        load_pk(req["circuit"], req["pk_path"], mt=req["mt"])
        return [prove(req["circuit"], a, w, mt=req["mt"])["proof_hex"] for a, w in req["jobs"]]
    raise ValueError(req["op"])


def _isolated(req, timeout=1800):
    import json
    import subprocess
    import sys
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    out = subprocess.run([sys.executable, "-c", _CHILD % os.path.dirname(HERE)], input=json.dumps(req), capture_output=True, text=True,
                         timeout=timeout, env=env)
    for line in out.stdout.splitlines():
        if line.startswith("RESULT "):
            return json.loads(line[7:])
    raise RuntimeError("reference child process failed: " + out.stderr[-2000:])


def verify_many(jobs):
    """jobs: (circuit, proof_hex, verify_args, vk_path or None).  The reference verifier (verify_<c>_proof ->
    r1cs_gg_ppzksnark_verifier_strong_IC) on explicit verification-key files, in a clean child process.  Returns a list of bools."""
    return _isolated({"op": "verify", "jobs": [(c, p, list(a), vk or os.path.join(KEY_DIR, c + "vk.txt")) for c, p, a, vk in jobs]})


def verify(circuit, proof_hex, verify_args, vk_path=None):
    """verify_args: the arguments of verify<Circuit>proof after the proof string."""
    return verify_many([(circuit, proof_hex, verify_args, vk_path)])[0]


def prove_isolated(circuit, jobs, pk_path=None, mt=True):
    """jobs: (gen<Circuit>proof args, pinned words).  Loads the pk with the reference's own parser and proves, in a clean child process.
    Returns the proof strings."""
    return _isolated({"op": "prove", "circuit": circuit, "pk_path": pk_path or os.path.join(KEY_DIR, circuit + "pk.txt"), "mt": mt,
                      "jobs": [(list(a), list(w)) for a, w in jobs]})
