"""ctypes bindings of include/zkb200.h (layer 2) and of the BlockMaze cgo surface (layer 1).

Wire formats: field elements 32-byte little-endian canonical; G1 affine 64 B (x y); G2 affine 128 B (x.c0 x.c1 y.c0 y.c1).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKB200_LIB") or os.path.join(_HERE, "libzkb200.so")
CIRCUITS = ("mint", "send", "deposit", "redeem")

if not os.path.exists(LIB_PATH):
    raise ImportError("blockmaze_b200: %s is missing -- run `python __graft_entry__.py` (nvcc, sm_100a). "
                      "There is no CPU fallback." % LIB_PATH)
lib = C.CDLL(LIB_PATH)

lib.zkb200_init.argtypes = [C.c_int]
lib.zkb200_last_error.restype = C.c_char_p
lib.zkb200_pk_load.restype = C.c_void_p
lib.zkb200_pk_load.argtypes = [C.c_char_p]
lib.zkb200_pk_free.argtypes = [C.c_void_p]
lib.zkb200_pk_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
lib.zkb200_prove.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_float)]
lib.zkb200_qap_witness_map.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_int)]
lib.zkb200_domain_op.restype = C.c_long
lib.zkb200_domain_op.argtypes = [C.c_size_t, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
lib.zkb200_msm_g1.argtypes = [C.c_size_t, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p]
lib.zkb200_msm_g2.argtypes = [C.c_size_t, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p]
lib.zkb200_field_op.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p]
lib.zkb200_bench_ntt.restype = C.c_float
lib.zkb200_bench_ntt.argtypes = [C.c_int, C.c_int, C.c_int]
lib.zkb200_bench_msm.restype = C.c_float
lib.zkb200_bench_msm.argtypes = [C.c_int, C.c_size_t, C.c_int, C.c_int]
lib.zkb200_last_breakdown_ms.argtypes = [C.POINTER(C.c_double)]
lib.zkb200_flush_l2.restype = None
lib.zkb200_device_sync.restype = None
lib.zkb200_device_timer.restype = C.c_float
lib.zkb200_device_timer.argtypes = [C.c_int]
lib.zkb200_bench_msm_slice.restype = C.c_float
lib.zkb200_bench_msm_slice.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_char_p]
lib.zkb200_bench_imad_peak.restype = C.c_float
lib.zkb200_bench_imad_peak.argtypes = [C.c_int]

lib.zkb200_set_key_dir.argtypes = [C.c_char_p]
lib.zkb200_set_random_words.argtypes = [C.c_void_p, C.c_size_t]
lib.zkb200_set_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
lib.zkb200_active_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
lib.zkb200_pk_load_on.restype = C.c_void_p
lib.zkb200_pk_load_on.argtypes = [C.c_char_p, C.c_int]
lib.zkb200_pk_device.argtypes = [C.c_void_p]
lib.zkb200_default_proof.restype = C.c_char_p
lib.zkb200_sched_create.restype = C.c_void_p
lib.zkb200_sched_create.argtypes = [C.c_int]
lib.zkb200_sched_pick.argtypes = [C.c_void_p]
lib.zkb200_sched_done.argtypes = [C.c_void_p, C.c_int]
lib.zkb200_sched_done.restype = None
lib.zkb200_sched_inflight.argtypes = [C.c_void_p, C.c_int]
lib.zkb200_sched_free.argtypes = [C.c_void_p]
lib.zkb200_sched_free.restype = None


class Tx(C.Structure):
    """zkb200_tx: one transaction of a zkb200_prove_batch call."""
    _fields_ = [("circuit", C.c_int), ("n", C.c_int), ("u", C.c_uint64 * 3), ("s", C.c_char_p * 14)]


lib.zkb200_prove_batch.argtypes = [C.c_size_t, C.POINTER(Tx), C.c_char_p, C.c_int]


class VTx(C.Structure):
    """zkb200_vtx: one proof of a zkb200_verify_batch call."""
    _fields_ = [("circuit", C.c_int), ("value_s", C.c_uint64), ("proof", C.c_char_p), ("s", C.c_char_p * 6)]


lib.zkb200_verify_batch.argtypes = [C.c_size_t, C.POINTER(VTx), C.c_char_p, C.c_int]

# gen<Circuit>proof argument types (SRC/<c>/<c>cgo.hpp) and verify<Circuit>proof argument types
GEN_SIGS = {
    "mint": [C.c_uint64, C.c_uint64] + [C.c_char_p] * 6 + [C.c_uint64, C.c_char_p],
    "redeem": [C.c_uint64, C.c_uint64] + [C.c_char_p] * 6 + [C.c_uint64, C.c_char_p],
    "send": [C.c_uint64] + [C.c_char_p] * 5 + [C.c_uint64, C.c_char_p, C.c_uint64] + [C.c_char_p] * 5,
    "deposit": [C.c_uint64, C.c_uint64] + [C.c_char_p] * 8 + [C.c_uint64] + [C.c_char_p] * 4 + [C.c_int, C.c_char_p, C.c_char_p],
}
VERIFY_SIGS = {"mint": [C.c_char_p] * 4 + [C.c_uint64], "redeem": [C.c_char_p] * 4 + [C.c_uint64], "send": [C.c_char_p] * 5,
               "deposit": [C.c_char_p] * 7}
NUM_VARS = {"mint": 151512, "send": 227046, "deposit": 457127, "redeem": 151579}
for _c in CIRCUITS:
    _g = getattr(lib, "gen%sproof" % _c.capitalize())
    _g.restype = C.c_void_p
    _g.argtypes = GEN_SIGS[_c]
    _v = getattr(lib, "verify%sproof" % _c.capitalize())
    _v.restype = C.c_bool
    _v.argtypes = VERIFY_SIGS[_c]
for _h, _sig in (("genCMT", [C.c_uint64, C.c_char_p, C.c_char_p]), ("computePRF", [C.c_char_p] * 2),
                 ("genCMTS", [C.c_uint64] + [C.c_char_p] * 3), ("computeCRH", [C.c_char_p] * 2), ("genRoot", [C.c_char_p, C.c_int])):
    getattr(lib, _h).restype = C.c_void_p
    getattr(lib, _h).argtypes = _sig
lib.zkb200_witness_mint.restype = C.c_long
lib.zkb200_witness_mint.argtypes = GEN_SIGS["mint"] + [C.c_int, C.c_void_p, C.c_size_t]
lib.zkb200_witness_send.restype = C.c_long
lib.zkb200_witness_send.argtypes = GEN_SIGS["send"] + [C.c_void_p, C.c_size_t]
lib.zkb200_witness_deposit.restype = C.c_long
lib.zkb200_witness_deposit.argtypes = GEN_SIGS["deposit"] + [C.c_void_p, C.c_size_t]

lib.zkb200_g1_sum.argtypes = [C.c_size_t, C.c_char_p, C.c_char_p]
lib.zkb200_witness_defer.argtypes = [C.c_int]
lib.zkb200_witness_defer.restype = None
lib.zkb200_last_assignment.restype = C.c_long
lib.zkb200_last_assignment.argtypes = [C.c_char_p, C.c_size_t]
lib.zkb200_g2_sum.argtypes = [C.c_size_t, C.c_char_p, C.c_char_p]
lib.zkb200_synth_scalars.argtypes = [C.c_size_t, C.c_size_t, C.c_char_p]
lib.zkb200_synth_bases.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_char_p]
lib.zkb200_set_isolate_h.argtypes = [C.c_int]
lib.zkb200_set_isolate_h.restype = None
lib.zkb200_pk_lanes.argtypes = [C.c_void_p]
lib.zkb200_lane_acquire.argtypes = [C.c_void_p]
lib.zkb200_lane_release.argtypes = [C.c_void_p, C.c_int]
lib.zkb200_lane_release.restype = None
lib.zkb200_prove_submit.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p]
lib.zkb200_prove_collect.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.c_float)]
lib.zkb200_keygen.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]

DOMAIN_OPS = {"FFT": 0, "iFFT": 1, "cosetFFT": 2, "icosetFFT": 3, "divide_by_Z_on_coset": 4}
FIELD_OPS = {"mul": 0, "add": 1, "sub": 2, "sqr": 3, "to_mont": 4, "from_mont": 5, "inverse": 6}


class ZkError(RuntimeError):
    pass


def last_error():
    return (lib.zkb200_last_error() or b"").decode()


def init(device=0):
    if lib.zkb200_init(device) != 0:
        raise ZkError(last_error())


def domain_size(min_size):
    kind = C.c_int(0)
    m = lib.zkb200_domain_op(min_size, 0, None, 0, C.byref(kind))
    if m < 0:
        raise ZkError(last_error())
    return m, ("step_radix2" if kind.value else "basic_radix2")


def domain_op(min_size, op, data):
    """data: bytes of m canonical 32-byte elements; returns the transformed bytes."""
    buf = C.create_string_buffer(bytes(data), len(data))
    m = lib.zkb200_domain_op(min_size, DOMAIN_OPS[op], buf, len(data) // 32, None)
    if m < 0:
        raise ZkError(last_error())
    return buf.raw


def msm_g1(bases, scalars, window_bits=0):
    out = C.create_string_buffer(64)
    if lib.zkb200_msm_g1(len(scalars) // 32, bytes(bases), bytes(scalars), window_bits, out) != 0:
        raise ZkError(last_error())
    return out.raw


def g1_sum(points):
    """Host-side sum of affine G1 points (64 B each): the last step of an MSM split by point range over several GPUs."""
    out = C.create_string_buffer(64)
    lib.zkb200_g1_sum(len(points), b"".join(points), out)
    return out.raw


def g2_sum(points):
    out = C.create_string_buffer(128)
    lib.zkb200_g2_sum(len(points), b"".join(points), out)
    return out.raw


def synth_scalars(first, n):
    """Scalars first .. first+n of the sweep's synthetic input: libff SHA512_rng<Fr>(i), 32 B canonical each."""
    out = C.create_string_buffer(32 * max(1, n))
    if lib.zkb200_synth_scalars(first, n, out) != 0:
        raise ZkError(last_error())
    return out.raw[:32 * n]


def synth_bases(group, first, n):
    """Bases first .. first+n of the sweep's synthetic input: SHA512_rng<Fr>(2^32 + i) * generator; group 1 = G1 (64 B), 2 = G2 (128 B)."""
    size = 64 * group
    out = C.create_string_buffer(size * max(1, n))
    if lib.zkb200_synth_bases(group, first, n, out) != 0:
        raise ZkError(last_error())
    return out.raw[:size * n]


def msm_g2(bases, scalars, window_bits=0):
    out = C.create_string_buffer(128)
    if lib.zkb200_msm_g2(len(scalars) // 32, bytes(bases), bytes(scalars), window_bits, out) != 0:
        raise ZkError(last_error())
    return out.raw


def field_op(field, op, a, b=None):
    """Raw Montgomery-representation field op on the GPU.  field: 'fr'|'fq'."""
    n = len(a) // 32
    out = C.create_string_buffer(32 * n)
    if lib.zkb200_field_op(0 if field == "fr" else 1, FIELD_OPS[op], n, bytes(a), bytes(b) if b is not None else None, out) != 0:
        raise ZkError(last_error())
    return out.raw


def _enc(args):
    return [a.encode() if isinstance(a, str) else a for a in args]


def set_key_dir(path):
    lib.zkb200_set_key_dir(os.fsencode(path))


def set_random_words(words):
    """Test hook: pin (r, s) with a std::random_device-style 32-bit word stream (empty list = back to the OS entropy source).
    Needs ZKB200_TEST_RNG=1 in the environment."""
    arr = (C.c_uint32 * max(1, len(words)))(*words)
    if lib.zkb200_set_random_words(C.cast(arr, C.c_void_p), len(words)) != 0:
        raise ZkError("zkb200_set_random_words refused: ZKB200_TEST_RNG=1 is not set")


def set_devices(devices=()):
    """Active devices of the cgo surface (gen_proof / prove_batch); empty = every visible device.  Returns how many."""
    arr = (C.c_int * max(1, len(devices)))(*devices)
    n = lib.zkb200_set_devices(arr, len(devices))
    if n < 0:
        raise ZkError("zkb200_set_devices: bad device list")
    return n


def active_devices():
    arr = (C.c_int * 64)()
    return list(arr[:lib.zkb200_active_devices(arr, 64)])


def _tx(circuit, args):
    """gen<Circuit>proof argument list -> zkb200_tx (uint64 and string arguments in order of appearance)."""
    t = Tx()
    t.circuit = CIRCUITS.index(circuit)
    ints = [a for a in args if isinstance(a, int)]
    strs = [a for a in args if not isinstance(a, int)]
    if circuit == "deposit":                      # the int `n` (number of leaves) sits among the strings
        sig = GEN_SIGS["deposit"]
        ints = [a for a, ty in zip(args, sig) if ty is C.c_uint64]
        t.n = [a for a, ty in zip(args, sig) if ty is C.c_int][0]
        strs = [a for a, ty in zip(args, sig) if ty is C.c_char_p]
    for i, v in enumerate(ints):
        t.u[i] = v
    for i, v in enumerate(strs):
        t.s[i] = v.encode() if isinstance(v, str) else v
    return t


def prove_batch(jobs, threads=0):
    """jobs: list of (circuit, gen<Circuit>proof argument list).  Proves them on every active device (zkb200_prove_batch).
    Returns (proof strings, number of default proofs)."""
    return prove_batch_run(prove_batch_prepare(jobs), threads)


def prove_batch_prepare(jobs):
    """The zkb200_tx array of a job list (ctypes marshalling of the argument strings: Python's cost, not the library's)."""
    arr = (Tx * max(1, len(jobs)))(*[_tx(c, a) for c, a in jobs])
    return arr, len(jobs)


def prove_batch_run(prepared, threads=0):
    arr, n = prepared
    out = C.create_string_buffer(513 * max(1, n))
    bad = lib.zkb200_prove_batch(n, arr, out, threads)
    if bad < 0:
        raise ZkError("zkb200_prove_batch: bad arguments")
    return [out.raw[513 * i:513 * i + 512].decode() for i in range(n)], bad


def helper(name, *args):
    """genCMT / computePRF / genCMTS / computeCRH / genRoot: returns the 64-char hex string."""
    return C.string_at(getattr(lib, name)(*_enc(args)), 64).decode()


def gen_proof(circuit, args):
    """gen<Circuit>proof through the BlockMaze cgo surface; returns the 512-char proof string."""
    p = getattr(lib, "gen%sproof" % circuit.capitalize())(*_enc(args))
    return C.string_at(p, 512).decode()


def last_breakdown_ms():
    """Host witness generation, zkb200_prove total, of which GPU, host finish -- of the last gen_proof call."""
    out = (C.c_double * 4)()
    lib.zkb200_last_breakdown_ms(out)
    return dict(zip(("witness", "prove", "gpu", "host_finish"), [round(float(x), 3) for x in out]))


def verify_proof(circuit, proof_hex, args):
    return bool(getattr(lib, "verify%sproof" % circuit.capitalize())(proof_hex.encode(), *_enc(args)))


def verify_batch(items, threads=0):
    """items: (circuit, proof_hex, verify<Circuit>proof arguments after the proof).  One random-linear-combination check for the whole
    batch (zkb200_verify_batch); returns the list of per-proof verdicts."""
    arr = (VTx * max(1, len(items)))()
    for t, (c, proof, args) in zip(arr, items):
        t.circuit = CIRCUITS.index(c)
        t.proof = proof.encode()
        strs = [a for a in args if not isinstance(a, int)]
        for i, v in enumerate(strs):
            t.s[i] = v.encode()
        ints = [a for a in args if isinstance(a, int)]
        t.value_s = ints[0] if ints else 0
    ok = C.create_string_buffer(max(1, len(items)))
    if lib.zkb200_verify_batch(len(items), arr, ok, threads) < 0:
        raise ZkError("zkb200_verify_batch: bad arguments or unreadable verification key")
    return [bool(b) for b in ok.raw[:len(items)]]


def verify_args(circuit, gen_args):
    """The verify<Circuit>proof arguments that go with a gen<Circuit>proof argument list."""
    a = gen_args
    if circuit in ("mint", "redeem"):
        return [a[6], a[2], a[7], a[8]]                       # cmtA_old, sn_old, cmtA, value_s
    if circuit == "send":
        return [a[5], a[2], a[4], a[11]]                      # cmtA_old, sn_old, cmtS, cmtA_new
    rt = "0x" + helper("genRoot", a[14], a[15])
    return [rt, a[11], a[8], a[2], a[9], a[6]]                # RT, pk, cmtB_old, sn_old, cmtB, sn_s


def witness(circuit, args, defer=False):
    """Full variable assignment (bytes, num_variables*32) computed by the native host generators.  defer=True: the SHA-256 gadget runs
    come from the seed expansion that gen_proof performs on the GPU (here: its host copy)."""
    lib.zkb200_witness_defer(1 if defer else 0)
    try:
        return _witness(circuit, args)
    finally:
        lib.zkb200_witness_defer(0)


def last_assignment(circuit):
    """The assignment on the GPU behind this thread's last gen_proof call (zkb200_last_assignment)."""
    n = NUM_VARS[circuit]
    out = C.create_string_buffer(32 * n)
    got = lib.zkb200_last_assignment(out, 32 * n)
    if got != n:
        raise ZkError("zkb200_last_assignment failed (%d)" % got)
    return out.raw


def _witness(circuit, args):
    n = NUM_VARS[circuit]
    out = C.create_string_buffer(32 * n)
    ptr = C.cast(out, C.c_void_p)
    if circuit in ("mint", "redeem"):
        got = lib.zkb200_witness_mint(*_enc(args), 1 if circuit == "redeem" else 0, ptr, n)
    elif circuit == "send":
        got = lib.zkb200_witness_send(*_enc(args), ptr, n)
    else:
        got = lib.zkb200_witness_deposit(*_enc(args), ptr, n)
    if got != n:
        raise ZkError("witness generation failed (%d)" % got)
    return out.raw


def keygen(cs_source_pk, out_pk, out_vk, words=()):
    """Generate a fresh key pair for the constraint system embedded in `cs_source_pk` (zkb200_keygen).  Returns phase seconds."""
    arr = (C.c_uint32 * max(1, len(words)))(*words)
    secs = (C.c_double * 3)()
    rc = lib.zkb200_keygen(os.fsencode(cs_source_pk), C.cast(arr, C.c_void_p), len(words), os.fsencode(out_pk), os.fsencode(out_vk), secs)
    if rc != 0:
        raise ZkError("zkb200_keygen failed (%d)" % rc)
    return [float(x) for x in secs]


class ProvingKey:
    """A proving key resident on the GPU (zkb200_pk_load)."""

    def __init__(self, path, device=None):
        self.handle = lib.zkb200_pk_load(os.fsencode(path)) if device is None else lib.zkb200_pk_load_on(os.fsencode(path), device)
        if not self.handle:
            raise ZkError(last_error())
        info = (C.c_uint64 * 8)()
        secs = (C.c_double * 3)()
        lib.zkb200_pk_info(self.handle, info, secs)
        (self.num_variables, self.num_inputs, self.num_constraints, self.domain_size, kind, self.nnz, self.num_coefficients,
         self.b_entries) = [int(x) for x in info]
        self.domain_kind = "step_radix2" if kind else "basic_radix2"
        self.load_seconds, self.parse_seconds, self.decompress_seconds = [float(x) for x in secs]

    def prove(self, assignment, r, s):
        """assignment: bytes (num_variables*32) or None to reuse the resident one; r, s: ints.
        Returns dict(rc, proof_hex, parts(384 B), timings_ms[gpu, qap, msm_h, host, msm_h_accumulate_kernel, A done, B done, L done])."""
        hexbuf = C.create_string_buffer(513)
        parts = C.create_string_buffer(384)
        tim = (C.c_float * 8)()
        if assignment is not None and len(assignment) != self.num_variables * 32:
            raise ValueError("assignment must be num_variables*32 bytes")
        rc = lib.zkb200_prove(self.handle, assignment, int(r).to_bytes(32, "little"), int(s).to_bytes(32, "little"), hexbuf, parts, tim)
        if rc < 0:
            raise ZkError(last_error())
        return dict(rc=rc, proof_hex=hexbuf.value.decode(), parts=parts.raw, timings_ms=list(tim), launches=lib.zkb200_last_launches())

    # ---- proofs in flight (zkb200.h "lanes"): submit enqueues the copy and all kernels, collect waits and assembles the proof
    @property
    def lanes(self):
        return int(lib.zkb200_pk_lanes(self.handle))

    def lane_acquire(self):
        return int(lib.zkb200_lane_acquire(self.handle))

    def lane_release(self, lane):
        lib.zkb200_lane_release(self.handle, lane)

    def submit(self, lane, assignment, r, s):
        """assignment: bytes or None to reuse what the lane holds.  The lane must be held (lane_acquire) and idle."""
        if assignment is not None and len(assignment) != self.num_variables * 32:
            raise ValueError("assignment must be num_variables*32 bytes")
        if lib.zkb200_prove_submit(self.handle, lane, assignment, int(r).to_bytes(32, "little"), int(s).to_bytes(32, "little")) != 0:
            raise ZkError("lane %d is not held or has a proof pending" % lane)

    def collect(self, lane, want_parts=False):
        hexbuf = C.create_string_buffer(513)
        parts = C.create_string_buffer(384) if want_parts else None
        tim = (C.c_float * 8)()
        rc = lib.zkb200_prove_collect(self.handle, lane, hexbuf, parts, tim)
        if rc < 0:
            raise ZkError("lane %d has nothing to collect" % lane)
        return dict(rc=rc, proof_hex=hexbuf.value.decode(), parts=parts.raw if parts else None, timings_ms=list(tim),
                    launches=lib.zkb200_last_launches())

    def qap_witness_map(self, assignment):
        out = C.create_string_buffer((self.domain_size + 1) * 32)
        sat = C.c_int(0)
        if lib.zkb200_qap_witness_map(self.handle, assignment, out, C.byref(sat)) != 0:
            raise ZkError(last_error())
        return out.raw, bool(sat.value)

    def close(self):
        if self.handle:
            lib.zkb200_pk_free(self.handle)
            self.handle = None

