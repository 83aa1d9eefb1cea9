"""GPU parity tests of the two hot kernels through the C-ABI: device field layer, evaluation-domain transforms (NTT),
multi-scalar multiplication.  Bit-exact against the oracle (small sizes), the reference itself (oracle/_ref, large sizes)
and size-independent properties at sweep sizes."""
import os
import random

import pytest

from oracle import bn254_oracle as O

pytestmark = pytest.mark.gpu


def fb(v):
    return b"".join(int(x).to_bytes(32, "little") for x in v)


def fl(b):
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def g1b(p):
    return bytes(64) if p is None else p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little")


@pytest.mark.parametrize("name,p", [("fr", O.R_MOD), ("fq", O.Q_MOD)])
def test_field_layer_montgomery_bit_exact(zk, name, p):
    """8x32-bit limb Montgomery arithmetic == Python integers on raw Montgomery representatives (fp.tcc:23-190 semantics)."""
    rng = random.Random(7)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 253), (1 << 254) % p, 0xFFFFFFFF, 1 << 32, (1 << 224) - 1, (1 << 256) % p]
    a = edge + [rng.randrange(p) for _ in range(20000)]
    b = list(reversed(edge)) + [rng.randrange(p) for _ in range(20000)]
    Ri = pow(1 << 256, -1, p)
    assert fl(zk.field_op(name, "mul", fb(a), fb(b))) == [x * y * Ri % p for x, y in zip(a, b)]
    assert fl(zk.field_op(name, "sqr", fb(a))) == [x * x * Ri % p for x in a]
    assert fl(zk.field_op(name, "add", fb(a), fb(b))) == [(x + y) % p for x, y in zip(a, b)]
    assert fl(zk.field_op(name, "sub", fb(a), fb(b))) == [(x - y) % p for x, y in zip(a, b)]
    assert fl(zk.field_op(name, "to_mont", fb(a))) == [x * (1 << 256) % p for x in a]
    assert fl(zk.field_op(name, "from_mont", fb(a))) == [x * Ri % p for x in a]
    inv = fl(zk.field_op(name, "inverse", fb(a[:3000])))                    # safegcd division steps on the device (ff.cuh inverse_gcd)
    assert inv == [(pow(x * Ri % p, -1, p) * (1 << 256) % p if x else 0) for x in a[:3000]]


def test_field_layer_against_reference(zk, ref):
    rng = random.Random(3)
    a = [rng.randrange(O.Q_MOD) for _ in range(512)]
    b = [rng.randrange(O.Q_MOD) for _ in range(512)]
    am, bm = ref.to_mont("fq", a), ref.to_mont("fq", b)
    prod = fl(zk.field_op("fq", "mul", fb(am), fb(bm)))
    assert prod == ref.to_mont("fq", ref.field_op("fq", "mul", a, b))


@pytest.mark.parametrize("ms", [2, 4, 8, 12, 24, 100, 768, 1024, 2048, 3000, 5000, 1 << 13, 3 << 12])
def test_domain_ops_against_oracle(zk, ms):
    """FFT/iFFT/cosetFFT/icosetFFT/divide_by_Z_on_coset for basic and step domains, incl. the rounding rule of the selector."""
    rng = random.Random(ms)
    dom = O.get_evaluation_domain(ms)
    assert zk.domain_size(ms) == (dom.m, dom.kind)
    v = [rng.randrange(O.R_MOD) for _ in range(dom.m)]
    v[0], v[-1] = 0, O.R_MOD - 1
    for op, f in (("FFT", dom.FFT), ("iFFT", dom.iFFT), ("cosetFFT", lambda x: dom.cosetFFT(x, 5)), ("icosetFFT", lambda x: dom.icosetFFT(x, 5)),
                  ("divide_by_Z_on_coset", dom.divide_by_Z_on_coset)):
        assert fl(zk.domain_op(ms, op, fb(v))) == f(v), (ms, op)


@pytest.mark.parametrize("ms", [1 << 16, 196608, 1 << 18, 167275, 1 << 19])
def test_domain_ops_against_reference_at_circuit_sizes(zk, ref, ms):
    """The four circuits' domains (mint/redeem: step 2^17+2^16; send 2^18; deposit 2^19) against libfqfft itself."""
    m, kind = zk.domain_size(ms)
    assert (m, kind) == ref.domain_size(ms)
    rng = random.Random(ms)
    raw = fb([rng.getrandbits(253) for _ in range(m)])
    for op in ("FFT", "iFFT", "cosetFFT", "icosetFFT", "divide_by_Z_on_coset"):
        assert zk.domain_op(ms, op, raw) == ref.domain_op_bytes(ms, op, raw), (ms, op)


@pytest.mark.parametrize("logn", [20, 22])
def test_ntt_properties_at_sweep_sizes(zk, logn):
    """Size-independent properties where the oracle is too slow: iFFT(FFT(a)) = a, icosetFFT(cosetFFT(a)) = a, linearity,
    and FFT of a delta = all-ones / of the constant polynomial = n*delta."""
    n = 1 << logn
    rng = random.Random(logn)
    raw = b"".join(rng.getrandbits(253).to_bytes(32, "little") for _ in range(n))
    ev = zk.domain_op(n, "FFT", raw)
    assert zk.domain_op(n, "iFFT", ev) == raw
    assert zk.domain_op(n, "icosetFFT", zk.domain_op(n, "cosetFFT", raw)) == raw
    delta = (1).to_bytes(32, "little") + bytes(32 * (n - 1))
    assert zk.domain_op(n, "FFT", delta) == (1).to_bytes(32, "little") * n
    ones = (1).to_bytes(32, "little") * n
    assert zk.domain_op(n, "iFFT", ones) == delta
    # spot-check against naive evaluation at a few domain points
    a = fl(raw[:32 * 64]) + [0] * 0
    sparse = raw[:32 * 64] + bytes(32 * (n - 64))
    evs = zk.domain_op(n, "FFT", sparse)
    w = O.get_root_of_unity(n)
    for i in (0, 1, 12345, n // 2 + 3, n - 1):
        x = pow(w, i, O.R_MOD)
        assert int.from_bytes(evs[32 * i:32 * i + 32], "little") == sum(c * pow(x, k, O.R_MOD) for k, c in enumerate(a)) % O.R_MOD


def _points(n):
    pts, P = [], O.G1_ONE
    for _ in range(n):
        P = O.G1.add(P, O.G1_ONE)
        pts.append(O.G1.to_affine(P))
    return pts


def test_msm_g1_edge_cases_against_oracle(zk):
    """Empty input, single point, zero / one scalars, r-1, repeated points (doubling inside a bucket), P and -P (cancellation
    to infinity), infinity among the bases -- the cases multi_exp_with_mixed_addition special-cases (multiexp.tcc:443-496)."""
    pts = _points(40)
    J = [O.G1.from_affine(p) for p in pts]
    inf = bytes(64)
    assert zk.msm_g1(b"", b"") == inf
    assert zk.msm_g1(g1b(pts[0]), fb([0])) == inf
    assert zk.msm_g1(g1b(pts[0]), fb([1])) == g1b(pts[0])
    assert zk.msm_g1(g1b(pts[3]), fb([O.R_MOD - 1])) == g1b(O.G1.to_affine(O.G1.neg(J[3])))
    neg = (pts[5][0], O.Q_MOD - pts[5][1])
    assert zk.msm_g1(g1b(pts[5]) + g1b(neg), fb([7, 7])) == inf
    assert zk.msm_g1(g1b(pts[5]) * 3, fb([5, 5, 5])) == g1b(O.G1.to_affine(O.G1.mul(15, J[5])))
    assert zk.msm_g1(g1b(pts[5]) + inf + g1b(pts[6]), fb([2, 99, 1])) == g1b(O.G1.to_affine(O.G1.add(O.G1.mul(2, J[5]), J[6])))
    rng = random.Random(2)
    for c in (0, 4, 7, 13, 16):
        sc = [rng.choice([0, 1, 2, rng.getrandbits(33), rng.randrange(O.R_MOD), O.R_MOD - 1, (1 << 253) + 5]) for _ in pts]
        exp = O.G1.to_affine(O.G1.multi_exp_with_mixed_addition(J, sc))
        assert zk.msm_g1(b"".join(map(g1b, pts)), fb(sc), c) == g1b(exp), c


@pytest.mark.parametrize("n", [1, 33, 1000, 1 << 14, 1 << 16])
def test_msm_g1_g2_against_reference(zk, ref, n):
    rng = random.Random(n)
    b1, b2 = ref.g1_bases_bytes(n, 12345), ref.g2_bases_bytes(n, 777)
    sc = [rng.randrange(O.R_MOD) for _ in range(n)]
    for i in range(0, n, 3):
        sc[i] = rng.choice([0, 1, rng.getrandbits(33), O.R_MOD - 1])
    s = fb(sc)
    assert zk.msm_g1(b1, s) == ref.msm_g1_bytes(b1, s, 1)[0]
    assert zk.msm_g2(b2, s) == ref.msm_g2_bytes(b2, s, 1)[0]


def test_msm_linearity_at_sweep_size(zk, ref):
    """MSM(s) + MSM(t) == MSM(s + t) and MSM(k*s) == k*MSM(s) at 2^18 points: a size the CPU reference takes seconds for."""
    n = 1 << 18
    rng = random.Random(1)
    b1 = ref.g1_bases_bytes(n, 31337)
    s = [rng.getrandbits(253) for _ in range(n)]
    t = [rng.getrandbits(253) for _ in range(n)]
    ps, pt = zk.msm_g1(b1, fb(s)), zk.msm_g1(b1, fb(t))
    pst = zk.msm_g1(b1, fb([(x + y) % O.R_MOD for x, y in zip(s, t)]))
    A, B = ref.g1_from(ps), ref.g1_from(pt)
    assert O.G1.to_affine(O.G1.add(O.G1.from_affine(A), O.G1.from_affine(B))) == ref.g1_from(pst)
    k = 0xDEADBEEFCAFE
    pk = zk.msm_g1(b1, fb([x * k % O.R_MOD for x in s]))
    assert ref.g1_mul(A, k) == ref.g1_from(pk)
    assert ps == ref.msm_g1_bytes(b1, fb(s), 0, chunks=0, mt=True)[0]


def test_msm_point_range_split_adds_up(zk):
    """A single MSM split by point range: the partial points of the slices add up to the full result."""
    import ctypes as C
    n, c = 50000, 12
    full, a, b = (C.create_string_buffer(64) for _ in range(3))
    zk.lib.zkb200_bench_msm_slice(1, 0, n, c, 1, full)
    zk.lib.zkb200_bench_msm_slice(1, 0, 20000, c, 1, a)
    zk.lib.zkb200_bench_msm_slice(1, 20000, n - 20000, c, 1, b)
    pt = lambda r: (int.from_bytes(r[:32], "little"), int.from_bytes(r[32:], "little"))
    s = O.G1.to_affine(O.G1.add(O.G1.from_affine(pt(a.raw)), O.G1.from_affine(pt(b.raw))))
    assert s == pt(full.raw) and O.G1.on_curve(s)
    # the fixed-base (expanded) layout gives the same point as the windowed one
    fb = C.create_string_buffer(64)
    zk.lib.zkb200_bench_msm_slice(1, 0, n, -16, 1, fb)
    assert fb.raw == full.raw


@pytest.mark.parametrize("n", [3000, 70000])
def test_msm_skewed_buckets_against_reference(zk, ref, n):
    """Bucket sizes as lopsided as a BlockMaze witness: most scalars share a handful of values, so a few buckets hold nearly all entries
    (the equal-range accumulation cuts them into many pieces and the CTA-wide fold sums them), the rest are empty or hold one entry."""
    rng = random.Random(n + 1)
    b1, b2 = ref.g1_bases_bytes(n, 99), ref.g2_bases_bytes(n, 98)
    big = (1 << 253) + 5
    sc = [rng.choices([2, 1, big, 0xFFFFFFFF, 0, None], [50, 20, 10, 10, 5, 5])[0] for _ in range(n)]
    sc = [rng.randrange(O.R_MOD) if x is None else x for x in sc]
    s = fb(sc)
    for c in (0, 8, 16):
        assert zk.msm_g1(b1, s, c) == ref.msm_g1_bytes(b1, s, 1)[0], c
    assert zk.msm_g2(b2, s) == ref.msm_g2_bytes(b2, s, 1)[0]
    # every scalar the same: ONE bucket per window holds all n entries
    s = fb([0x1234567] * n)
    assert zk.msm_g1(b1, s) == ref.msm_g1_bytes(b1, s, 1)[0]
    assert zk.g1_sum([zk.msm_g1(b1[:64 * 1000], s[:32 * 1000]), zk.msm_g1(b1[64 * 1000:], s[32 * 1000:])]) == zk.msm_g1(b1, s)


def test_synthetic_sweep_inputs_are_the_documented_streams(zk):
    """zkb200_synth_scalars / zkb200_synth_bases: scalar i == libff SHA512_rng<Fr>(i) (oracle restatement, pinned against libff in
    tests/test_oracle_pinned.py); base i == SHA512_rng(2^32 + i) * generator for G1 and G2.  Slices are functions of the global index."""
    n = 3000
    want = [O.sha512_rng(i) for i in range(n)]
    assert fl(zk.synth_scalars(0, n)) == want
    assert fl(zk.synth_scalars(1000, 50)) == want[1000:1050]
    assert fl(zk.synth_scalars((1 << 33) + 5, 4)) == [O.sha512_rng((1 << 33) + 5 + i) for i in range(4)]
    b1 = zk.synth_bases(1, 0, 40)
    for i in (0, 1, 39):
        p = O.G1.to_affine(O.G1.mul(O.sha512_rng((1 << 32) + i), O.G1_ONE))
        assert b1[64 * i:64 * i + 64] == g1b(p)
    assert zk.synth_bases(1, 17, 5) == b1[64 * 17:64 * 22]
    b2 = zk.synth_bases(2, 0, 6)
    assert zk.synth_bases(2, 3, 2) == b2[128 * 3:128 * 5]


def test_synthetic_g2_bases_against_reference(zk, ref):
    b2 = zk.synth_bases(2, 5, 3)
    gen = ref.g2_from(_ref_g2_gen(ref))
    for i in range(3):
        assert ref.g2_bytes(ref.g2_mul(gen, O.sha512_rng((1 << 32) + 5 + i))) == b2[128 * i:128 * i + 128]


def _ref_g2_gen(ref):
    import ctypes as C
    out = C.create_string_buffer(128)
    ref.lib("kernels").ref_g2_gen(out)
    return out.raw


@pytest.mark.slow
def test_sweep_sizes_against_reference_on_identical_inputs(zk, ref):
    """the host legs of the sweep see the inputs the GPU sees, and the results are compared -- G1 MSM at 2^20 and G2 MSM
    at 2^18 against libff multi_exp (all host threads), NTT at 2^20 element-wise against libfqfft, on the SHA512_rng streams."""
    import ctypes as C
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    n = 1 << 20
    sc = zk.synth_scalars(0, n)
    b1 = zk.synth_bases(1, 0, n)
    got = C.create_string_buffer(64)
    zk.lib.zkb200_bench_msm_slice(1, 0, n, 0, 1, got)                  # the bench's own device-resident path
    want = ref.msm_g1_bytes(b1, sc, 0, chunks=0, mt=True)[0]
    assert got.raw == want
    assert zk.msm_g1(b1, sc) == want                                     # and the host-buffer entry point
    fx = C.create_string_buffer(64)
    zk.lib.zkb200_bench_msm_slice(1, 0, n, -16, 1, fx)                 # fixed-base layout (what the prover's H query uses)
    assert fx.raw == want
    m = 1 << 18
    b2 = zk.synth_bases(2, 0, m)
    got2 = C.create_string_buffer(128)
    zk.lib.zkb200_bench_msm_slice(2, 0, m, 0, 1, got2)
    assert got2.raw == ref.msm_g2_bytes(b2, sc[:32 * m], 0, chunks=0, mt=True)[0]
    assert zk.domain_op(n, "FFT", sc) == ref.domain_op_bytes(n, "FFT", sc, mt=True)
    # a split by point range adds up to the same points (G1 on 8 slices, G2 on 3)
    parts = []
    for k in range(8):
        o = C.create_string_buffer(64)
        zk.lib.zkb200_bench_msm_slice(1, k * (n // 8), n // 8, 0, 1, o)
        parts.append(o.raw)
    assert zk.g1_sum(parts) == want
    parts, cuts = [], [0, 100000, 100001, m]
    for a, b in zip(cuts, cuts[1:]):
        o = C.create_string_buffer(128)
        zk.lib.zkb200_bench_msm_slice(2, a, b - a, 0, 1, o)
        parts.append(o.raw)
    assert zk.g2_sum(parts) == got2.raw


@pytest.mark.slow
def test_ntt_round_trip_at_2_24(zk):
    """The largest sweep size (512 MB per vector, three passes): iFFT(FFT(a)) == a and icosetFFT(cosetFFT(a)) == a on the SHA512_rng stream,
    plus a few evaluations checked against Horner on the host."""
    n = 1 << 24
    raw = zk.synth_scalars(0, n)
    ev = zk.domain_op(n, "FFT", raw)
    assert zk.domain_op(n, "iFFT", ev) == raw
    del ev
    assert zk.domain_op(n, "icosetFFT", zk.domain_op(n, "cosetFFT", raw)) == raw
    k = 48
    sparse = raw[:32 * k] + bytes(32 * (n - k))
    evs = zk.domain_op(n, "FFT", sparse)
    a = fl(raw[:32 * k])
    w = O.get_root_of_unity(n)
    for i in (0, 1, 54321, n // 2 + 7, n - 1):
        x = pow(w, i, O.R_MOD)
        assert int.from_bytes(evs[32 * i:32 * i + 32], "little") == sum(c * pow(x, j, O.R_MOD) for j, c in enumerate(a)) % O.R_MOD
