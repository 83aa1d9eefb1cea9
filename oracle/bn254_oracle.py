"""TEST INFRASTRUCTURE -- CPU restatement (oracle) of BlockMaze's prover hot path.  NOT product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.  It restates, with
Python integers, the algorithms of the reference (paths relative to /root/reference/libsnark-vnt/):

  * Fr / Fq Montgomery representation          depends/libsnark/depends/libff/libff/algebra/fields/fp.tcc:23-190
  * Fr::random_element word consumption        .../fields/fp.tcc:695-721, .../fields/bigint.tcc:167-179
  * get_root_of_unity                          .../fields/field_utils.tcc:36-51
  * get_evaluation_domain                      depends/libsnark/depends/libfqfft/libfqfft/evaluation_domain/get_evaluation_domain.tcc:33-52
  * basic_radix2_domain                        .../evaluation_domain/domains/basic_radix2_domain.tcc:26-112, basic_radix2_domain_aux.tcc:45-79,172-180
  * step_radix2_domain                         .../evaluation_domain/domains/step_radix2_domain.tcc:21-248
  * r1cs_to_qap_witness_map                    depends/libsnark/libsnark/reductions/r1cs_to_qap/r1cs_to_qap.tcc:205-334
  * alt_bn128 G1/G2 group law                  .../libff/algebra/curves/alt_bn128/alt_bn128_g1.cpp:139-358, alt_bn128_g2.cpp:149-365
  * multi_exp (BDLO12) and the 0/1 prefilter   .../libff/algebra/scalar_multiplication/multiexp.tcc:165-282,443-496
  * kc_multi_exp_with_mixed_addition           depends/libsnark/libsnark/knowledge_commitment/kc_multiexp.tcc:21-89
  * r1cs_gg_ppzksnark_prover                   depends/libsnark/libsnark/zk_proof_systems/ppzksnark/r1cs_gg_ppzksnark/r1cs_gg_ppzksnark.tcc:390-506
  * pk / vk file grammar                       .../r1cs_gg_ppzksnark.tcc:52-108 (+ SURVEY.md Appendix A)
  * proof hex encoding                         src/mint/mintcgo.cpp:112-187
  * uint256S / Note::cm / Compute_PRF / CRH / incremental Merkle tree (depth 8)
                                               src/mint/uint256.h:200-226, src/deposit/Note.h:30-78, src/deposit/util.h:231-258,
                                               src/deposit/IncrementalMerkleTree.hpp

PARITY PINNING: this restatement is checked against the reference itself (compiled unmodified into oracle/_ref by
oracle/Makefile) in tests/test_oracle_pinned.py, and against golden vectors generated from the reference and
committed under tests/golden/ (generator: tests/golden/make_golden.py).
"""
import hashlib

# ---------------------------------------------------------------------------------------------------------------
# constants (alt_bn128_init.cpp:96-144, 190-206, 261-269)
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617   # Fr modulus r
Q_MOD = 21888242871839275222246405745257275088696311157297823662689037894645226208583   # Fq modulus q
MONT_R = 1 << 256
FR_S = 28
FR_ROOT_OF_UNITY = 19103219067921713944291392827692070036145651957329286315305642004821462161904  # 2^28-th root
FR_GENERATOR = 5                                                                            # multiplicative_generator
G1_B = 3
G1_GEN = (1, 2)
# twist coefficient b' = 3 / (9 + u) in Fq2 = Fq[u]/(u^2+1)
G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))


def fr_to_mont(x): return x * MONT_R % R_MOD
def fr_from_mont(x): return x * pow(MONT_R, -1, R_MOD) % R_MOD
def fq_to_mont(x): return x * MONT_R % Q_MOD
def fq_from_mont(x): return x * pow(MONT_R, -1, Q_MOD) % Q_MOD


def ceil_log2(n):
    """libff::log2 (common/utils.cpp:32-45): ceil(log2(n))."""
    r = 0 if (n & (n - 1)) == 0 else 1
    while n > 1:
        n >>= 1
        r += 1
    return r


def fr_from_words(words, count=2):
    """Fr::random_element fed by a std::random_device word stream (fp.tcc:695-721, bigint.tcc:167-179).
    8 x 32-bit words -> mont_repr (word 0 = least significant), bits >= 254 cleared, rejected if >= r.
    The field element is the value whose MONTGOMERY representation is that integer.  Returns canonical values."""
    out, pos = [], 0
    while len(out) < count:
        if pos + 8 > len(words):
            raise ValueError("random word stream exhausted")
        v = 0
        for i in range(8):
            v |= (words[pos + i] & 0xFFFFFFFF) << (32 * i)
        pos += 8
        v &= (1 << 254) - 1
        if v < R_MOD:
            out.append(fr_from_mont(v))
    return out


def fixed_rng_words(seed, n):
    """The word stream of oracle/ref_harness/fixed_rng.cpp (splitmix64 counter, low 32 bits)."""
    M = (1 << 64) - 1
    out = []
    for ctr in range(1, n + 1):
        z = (seed + ctr * 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        out.append(z & 0xFFFFFFFF)
    return out


# ---------------------------------------------------------------------------------------------------------------
# evaluation domains
def get_root_of_unity(n):
    logn = ceil_log2(n)
    if n != 1 << logn or logn > FR_S:
        raise ValueError("no root of unity of that order")
    w = FR_ROOT_OF_UNITY
    for _ in range(FR_S - logn):
        w = w * w % R_MOD
    return w


def bitreverse(k, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (k & 1)
        k >>= 1
    return r


def basic_radix2_fft(a, omega):
    """_basic_serial_radix2_FFT (basic_radix2_domain_aux.tcc:45-79): CLRS in-place, bit-reversal first."""
    p = R_MOD
    n = len(a)
    logn = ceil_log2(n)
    assert n == 1 << logn
    a = list(a)
    for k in range(n):
        rk = bitreverse(k, logn)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for _ in range(logn):
        w_m = pow(omega, n // (2 * m), p)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = w * a[k + j + m] % p
                a[k + j + m] = (a[k + j] - t) % p
                a[k + j] = (a[k + j] + t) % p
                w = w * w_m % p
        m *= 2
    return a


def multiply_by_coset(a, g):
    """_multiply_by_coset (basic_radix2_domain_aux.tcc:172-180): a[i] *= g^i."""
    out, u = [a[0]] if a else [], g
    for i in range(1, len(a)):
        out.append(a[i] * u % R_MOD)
        u = u * g % R_MOD
    return out


class BasicRadix2Domain:
    kind = "basic_radix2"

    def __init__(self, m):
        if m <= 1:
            raise ValueError("basic_radix2(): expected m > 1")
        logm = ceil_log2(m)
        if m != 1 << logm or logm > FR_S:
            raise ValueError("basic_radix2(): bad size")
        self.m = m
        self.omega = get_root_of_unity(m)

    def FFT(self, a):
        assert len(a) == self.m
        return basic_radix2_fft(a, self.omega)

    def iFFT(self, a):
        assert len(a) == self.m
        inv = pow(self.m, -1, R_MOD)
        return [x * inv % R_MOD for x in basic_radix2_fft(a, pow(self.omega, -1, R_MOD))]

    def cosetFFT(self, a, g): return self.FFT(multiply_by_coset(a, g))
    def icosetFFT(self, a, g): return multiply_by_coset(self.iFFT(a), pow(g, -1, R_MOD))
    def get_domain_element(self, idx): return pow(self.omega, idx, R_MOD)
    def compute_vanishing_polynomial(self, t): return (pow(t, self.m, R_MOD) - 1) % R_MOD

    def divide_by_Z_on_coset(self, P):
        zinv = pow(self.compute_vanishing_polynomial(FR_GENERATOR), -1, R_MOD)
        return [x * zinv % R_MOD for x in P]


class StepRadix2Domain:
    kind = "step_radix2"

    def __init__(self, m):
        if m <= 1:
            raise ValueError("step_radix2(): expected m > 1")
        self.m = m
        self.big_m = 1 << (ceil_log2(m) - 1)
        self.small_m = m - self.big_m
        if self.small_m != 1 << ceil_log2(self.small_m):
            raise ValueError("step_radix2(): expected small_m == 1<<log2(small_m)")
        self.omega = get_root_of_unity(1 << ceil_log2(m))
        self.big_omega = self.omega * self.omega % R_MOD
        self.small_omega = get_root_of_unity(self.small_m)

    def FFT(self, a):
        p, big_m, small_m = R_MOD, self.big_m, self.small_m
        assert len(a) == self.m
        c, d = [0] * big_m, [0] * big_m
        w = 1
        for i in range(big_m):
            if i < small_m:
                c[i] = (a[i] + a[i + big_m]) % p
                d[i] = w * (a[i] - a[i + big_m]) % p
            else:
                c[i] = a[i]
                d[i] = w * a[i] % p
            w = w * self.omega % p
        compr = 1 << (ceil_log2(big_m) - ceil_log2(small_m))
        e = [0] * small_m
        for i in range(small_m):
            e[i] = sum(d[i + j * small_m] for j in range(compr)) % p
        c = basic_radix2_fft(c, self.big_omega)
        e = basic_radix2_fft(e, self.small_omega)
        return c + e

    def iFFT(self, a):
        p, big_m, small_m = R_MOD, self.big_m, self.small_m
        assert len(a) == self.m
        U0 = basic_radix2_fft(a[:big_m], pow(self.big_omega, -1, p))
        U1 = basic_radix2_fft(a[big_m:], pow(self.small_omega, -1, p))
        i0, i1 = pow(big_m, -1, p), pow(small_m, -1, p)
        U0 = [x * i0 % p for x in U0]
        U1 = [x * i1 % p for x in U1]
        tmp, w = [], 1
        for i in range(big_m):
            tmp.append(U0[i] * w % p)
            w = w * self.omega % p
        out = [0] * self.m
        for i in range(small_m, big_m):
            out[i] = U0[i]
        compr = 1 << (ceil_log2(big_m) - ceil_log2(small_m))
        for i in range(small_m):
            for j in range(1, compr):
                U1[i] = (U1[i] - tmp[i + j * small_m]) % p
        winv, w = pow(self.omega, -1, p), 1
        for i in range(small_m):
            U1[i] = U1[i] * w % p
            w = w * winv % p
        half = pow(2, -1, p)
        for i in range(small_m):
            out[i] = (U0[i] + U1[i]) * half % p
            out[big_m + i] = (U0[i] - U1[i]) * half % p
        return out

    def cosetFFT(self, a, g): return self.FFT(multiply_by_coset(a, g))
    def icosetFFT(self, a, g): return multiply_by_coset(self.iFFT(a), pow(g, -1, R_MOD))

    def get_domain_element(self, idx):
        if idx < self.big_m:
            return pow(self.big_omega, idx, R_MOD)
        return self.omega * pow(self.small_omega, idx - self.big_m, R_MOD) % R_MOD

    def compute_vanishing_polynomial(self, t):
        p = R_MOD
        return (pow(t, self.big_m, p) - 1) * (pow(t, self.small_m, p) - pow(self.omega, self.small_m, p)) % p

    def divide_by_Z_on_coset(self, P):
        p, big_m, small_m, g, om = R_MOD, self.big_m, self.small_m, FR_GENERATOR, self.omega
        Z0 = (pow(g, big_m, p) - 1) % p
        c_sm_Z0 = pow(g, small_m, p) * Z0 % p
        o_sm_Z0 = pow(om, small_m, p) * Z0 % p
        o_2sm = pow(om, 2 * small_m, p)
        out, elt = [], 1
        for i in range(big_m):
            out.append(P[i] * pow((c_sm_Z0 * elt - o_sm_Z0) % p, -1, p) % p)
            elt = elt * o_2sm % p
        go = g * om % p
        Z1 = (pow(go, big_m, p) - 1) * (pow(go, small_m, p) - pow(om, small_m, p)) % p
        z1inv = pow(Z1, -1, p)
        for i in range(small_m):
            out.append(P[big_m + i] * z1inv % p)
        return out


def get_evaluation_domain(min_size):
    """get_evaluation_domain.tcc:33-52 restricted to the domains reachable for BN254 Fr (extended_radix2 only applies
    when logm == s+1 = 29; geometric/arithmetic need Fr::geometric_generator which alt_bn128 does not define)."""
    big = 1 << (ceil_log2(min_size) - 1)
    small = min_size - big
    rounded_small = 1 << ceil_log2(small)
    for ctor, size in ((BasicRadix2Domain, min_size), (StepRadix2Domain, min_size),
                       (BasicRadix2Domain, big + rounded_small), (StepRadix2Domain, big + rounded_small)):
        try:
            return ctor(size)
        except ValueError:
            continue
    raise ValueError("get_evaluation_domain: no matching domain")


# ---------------------------------------------------------------------------------------------------------------
# R1CS -> QAP witness map
class R1CS:
    """rows: list of (a, b, c), each a list of (var_index, coeff) with index 0 = constant ONE."""

    def __init__(self, num_inputs, num_aux, rows):
        self.num_inputs, self.num_aux, self.rows = num_inputs, num_aux, rows

    @property
    def num_variables(self): return self.num_inputs + self.num_aux
    @property
    def num_constraints(self): return len(self.rows)


def lc_eval(terms, full):
    acc = 0
    for idx, coeff in terms:
        acc += coeff * (1 if idx == 0 else full[idx - 1])
    return acc % R_MOD


def is_satisfied(cs, full):
    return all(lc_eval(a, full) * lc_eval(b, full) % R_MOD == lc_eval(c, full) for a, b, c in cs.rows)


def r1cs_to_qap_witness_map(cs, full, return_intermediates=False):
    """r1cs_to_qap.tcc:205-334 with d1=d2=d3=0.  Returns coefficients_for_H (m+1 canonical values)."""
    p = R_MOD
    dom = get_evaluation_domain(cs.num_constraints + cs.num_inputs + 1)
    m, nc = dom.m, cs.num_constraints
    aA, aB, aC = [0] * m, [0] * m, [0] * m
    for i in range(cs.num_inputs + 1):
        aA[i + nc] = 1 if i == 0 else full[i - 1]
    for i, (a, b, c) in enumerate(cs.rows):
        aA[i] = (aA[i] + lc_eval(a, full)) % p
        aB[i] = (aB[i] + lc_eval(b, full)) % p
        aC[i] = (aC[i] + lc_eval(c, full)) % p
    evalA, evalB, evalC = list(aA), list(aB), list(aC)
    aA, aB = dom.iFFT(aA), dom.iFFT(aB)
    aA, aB = dom.cosetFFT(aA, FR_GENERATOR), dom.cosetFFT(aB, FR_GENERATOR)
    H = [x * y % p for x, y in zip(aA, aB)]
    aC = dom.cosetFFT(dom.iFFT(aC), FR_GENERATOR)
    H = [(x - y) % p for x, y in zip(H, aC)]
    H = dom.divide_by_Z_on_coset(H)
    H = dom.icosetFFT(H, FR_GENERATOR)
    coeffs = H + [0]
    if return_intermediates:
        return coeffs, dict(domain=dom, evalA=evalA, evalB=evalB, evalC=evalC)
    return coeffs


# ---------------------------------------------------------------------------------------------------------------
# Fq2 and the curve groups.  Points are Jacobian (X, Y, Z) with Z == 0 for infinity; affine (x, y) or None.
class Fq2:
    __slots__ = ("c0", "c1")

    def __init__(self, c0, c1=0): self.c0, self.c1 = c0 % Q_MOD, c1 % Q_MOD
    def __add__(s, o): return Fq2(s.c0 + o.c0, s.c1 + o.c1)
    def __sub__(s, o): return Fq2(s.c0 - o.c0, s.c1 - o.c1)
    def __neg__(s): return Fq2(-s.c0, -s.c1)
    def __eq__(s, o): return s.c0 == o.c0 and s.c1 == o.c1
    def __hash__(s): return hash((s.c0, s.c1))

    def __mul__(s, o):
        if isinstance(o, int):
            return Fq2(s.c0 * o, s.c1 * o)
        return Fq2(s.c0 * o.c0 - s.c1 * o.c1, s.c0 * o.c1 + s.c1 * o.c0)   # u^2 = -1 (alt_bn128_init.cpp:151)

    def inv(s):
        d = pow(s.c0 * s.c0 + s.c1 * s.c1, -1, Q_MOD)
        return Fq2(s.c0 * d, -s.c1 * d)

    def is_zero(s): return s.c0 == 0 and s.c1 == 0
    def __repr__(s): return "Fq2(%d, %d)" % (s.c0, s.c1)


class _FqOps:
    zero, one = 0, 1
    @staticmethod
    def add(a, b): return (a + b) % Q_MOD
    @staticmethod
    def sub(a, b): return (a - b) % Q_MOD
    @staticmethod
    def mul(a, b): return a * b % Q_MOD
    @staticmethod
    def inv(a): return pow(a, -1, Q_MOD)
    @staticmethod
    def is_zero(a): return a % Q_MOD == 0


class _Fq2Ops:
    zero, one = Fq2(0), Fq2(1)
    @staticmethod
    def add(a, b): return a + b
    @staticmethod
    def sub(a, b): return a - b
    @staticmethod
    def mul(a, b): return a * b
    @staticmethod
    def inv(a): return a.inv()
    @staticmethod
    def is_zero(a): return a.is_zero()


TWIST_B = Fq2(3) * Fq2(9, 1).inv()      # alt_bn128_init.cpp:190-192


class Group:
    def __init__(self, F, b):
        self.F, self.b = F, b

    def zero(self): return (self.F.zero, self.F.one, self.F.zero)
    def is_zero(self, P): return self.F.is_zero(P[2])
    def from_affine(self, a): return self.zero() if a is None else (a[0], a[1], self.F.one)

    def to_affine(self, P):
        F = self.F
        if self.is_zero(P):
            return None
        zi = F.inv(P[2])
        zi2 = F.mul(zi, zi)
        return (F.mul(P[0], zi2), F.mul(P[1], F.mul(zi2, zi)))

    def on_curve(self, a):
        F = self.F
        if a is None:
            return True
        x, y = a
        return F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), self.b)) == F.zero

    def dbl(self, P):
        F = self.F
        if self.is_zero(P):
            return P
        X, Y, Z = P
        A, B = F.mul(X, X), F.mul(Y, Y)
        C = F.mul(B, B)
        t = F.add(X, B)
        D = F.sub(F.sub(F.mul(t, t), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fv = F.mul(E, E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        YZ = F.mul(Y, Z)
        return (X3, Y3, F.add(YZ, YZ))

    def add(self, P, Q):
        F = self.F
        if self.is_zero(P):
            return Q
        if self.is_zero(Q):
            return P
        X1, Y1, Z1 = P
        X2, Y2, Z2 = Q
        Z1Z1, Z2Z2 = F.mul(Z1, Z1), F.mul(Z2, Z2)
        U1, U2 = F.mul(X1, Z2Z2), F.mul(X2, Z1Z1)
        S1, S2 = F.mul(Y1, F.mul(Z2, Z2Z2)), F.mul(Y2, F.mul(Z1, Z1Z1))
        if U1 == U2:
            if S1 == S2:
                return self.dbl(P)
            return self.zero()
        H = F.sub(U2, U1)
        Rr = F.sub(S2, S1)
        HH = F.mul(H, H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.mul(Rr, Rr), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(Rr, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(Z1, Z2), H)
        return (X3, Y3, Z3)

    def neg(self, P): return (P[0], self.F.sub(self.F.zero, P[1]), P[2])

    def mul(self, k, P):
        R, k = self.zero(), k
        for bit in bin(k)[2:] if k else "":
            R = self.dbl(R)
            if bit == "1":
                R = self.add(R, P)
        return R

    # ---- multi-exponentiation, structured as the reference's BDLO12 (multiexp.tcc:165-282) -------------------
    def multi_exp_inner(self, bases, scalars):
        length = len(bases)
        if length == 0:
            return self.zero()
        L = ceil_log2(length)
        c = L - (L // 3 - 2)
        num_bits = max(s.bit_length() for s in scalars)
        num_groups = (num_bits + c - 1) // c
        result = self.zero()
        for k in range(num_groups - 1, -1, -1):
            for _ in range(c):
                result = self.dbl(result)
            buckets = {}
            for B, s in zip(bases, scalars):
                idx = (s >> (k * c)) & ((1 << c) - 1)
                if idx:
                    buckets[idx] = self.add(buckets[idx], B) if idx in buckets else B
            # sum_i i*bucket[i] via running sums over the occupied ids (same value as the dense suffix-sum loop)
            running, prev = self.zero(), None
            acc = self.zero()
            for idx in sorted(buckets, reverse=True):
                if prev is not None:
                    acc = self.add(acc, self.mul(prev - idx, running))
                running = self.add(running, buckets[idx])
                prev = idx
            if prev is not None:
                acc = self.add(acc, self.mul(prev, running))
            result = self.add(result, acc)
        return result

    def multi_exp_with_mixed_addition(self, bases, scalars):
        """multiexp.tcc:443-496: scalar 0 skipped, scalar 1 added directly, the rest through BDLO12."""
        acc, rb, rs = self.zero(), [], []
        for B, s in zip(bases, scalars):
            if s == 0:
                continue
            if s == 1:
                acc = self.add(acc, B)
            else:
                rb.append(B)
                rs.append(s)
        return self.add(acc, self.multi_exp_inner(rb, rs))


G1 = Group(_FqOps, 3)
G2 = Group(_Fq2Ops, TWIST_B)
G1_ONE = (1, 2, 1)
G2_ONE = (Fq2(*G2_GEN[0]), Fq2(*G2_GEN[1]), Fq2(1))


# ---------------------------------------------------------------------------------------------------------------
# proving key / verification key files (hybrid text + binary)
def fq_sqrt(a):
    """q = 3 mod 4."""
    y = pow(a, (Q_MOD + 1) // 4, Q_MOD)
    if y * y % Q_MOD != a % Q_MOD:
        raise ValueError("not a square")
    return y


def fq2_sqrt(a):
    """Any square root in Fq2 = Fq[u]/(u^2+1) (complex method); the caller fixes the sign."""
    if a.c1 == 0:
        try:
            return Fq2(fq_sqrt(a.c0), 0)
        except ValueError:
            return Fq2(0, fq_sqrt(-a.c0 % Q_MOD))
    n = fq_sqrt((a.c0 * a.c0 + a.c1 * a.c1) % Q_MOD)
    half = pow(2, -1, Q_MOD)
    for cand in ((a.c0 + n) * half % Q_MOD, (a.c0 - n) * half % Q_MOD):
        try:
            x0 = fq_sqrt(cand)
        except ValueError:
            continue
        if x0 == 0:
            continue
        x1 = a.c1 * pow(2 * x0, -1, Q_MOD) % Q_MOD
        r = Fq2(x0, x1)
        if r * r == a:
            return r
    raise ValueError("not a square in Fq2")


class _Reader:
    def __init__(self, data): self.d, self.pos = data, 0

    def dec(self):
        """ASCII decimal terminated by one whitespace byte."""
        e = self.pos
        while self.d[e] not in b" \n":
            e += 1
        v = int(self.d[self.pos:e])
        self.pos = e + 1
        return v

    def expect(self, b):
        assert self.d[self.pos:self.pos + len(b)] == b, (self.pos, self.d[self.pos:self.pos + 16])
        self.pos += len(b)

    def g1(self):
        """alt_bn128_g1.cpp:404-465: '0'|'1' zero flag, 32-byte LE Montgomery x, '0'|'1' = LSB of canonical y."""
        z = self.d[self.pos:self.pos + 1]
        xm = int.from_bytes(self.d[self.pos + 1:self.pos + 33], "little")
        lsb = self.d[self.pos + 33] - 48
        self.pos += 34
        if z == b"1":
            return None
        x = fq_from_mont(xm)
        y = fq_sqrt((x * x * x + 3) % Q_MOD)
        if (y & 1) != lsb:
            y = Q_MOD - y
        return (x, y)

    def g2(self):
        """alt_bn128_g2.cpp:418-478: zero flag, x.c0, x.c1 (32-byte LE Montgomery), LSB of canonical y.c0."""
        z = self.d[self.pos:self.pos + 1]
        x0 = int.from_bytes(self.d[self.pos + 1:self.pos + 33], "little")
        x1 = int.from_bytes(self.d[self.pos + 33:self.pos + 65], "little")
        lsb = self.d[self.pos + 65] - 48
        self.pos += 66
        if z == b"1":
            return None
        x = Fq2(fq_from_mont(x0), fq_from_mont(x1))
        y = fq2_sqrt(x * x * x + TWIST_B)
        if (y.c0 & 1) != lsb:
            y = -y
        return (x, y)


class ProvingKey:
    pass


def parse_pk(data, max_points=None):
    """Parse a reference pk file.  max_points (testing aid) limits how many points of each query are decompressed
    (the rest are skipped byte-wise); the embedded constraint system is always parsed in full."""
    rd = _Reader(data)
    pk = ProvingKey()
    pk.alpha_g1 = rd.g1(); rd.expect(b"\n")
    pk.beta_g1 = rd.g1(); rd.expect(b"\n")
    pk.beta_g2 = rd.g2(); rd.expect(b"\n")
    pk.delta_g1 = rd.g1(); rd.expect(b"\n")
    pk.delta_g2 = rd.g2(); rd.expect(b"\n")

    def g1_vec():
        n = rd.dec()
        out = []
        for i in range(n):
            if max_points is not None and i >= max_points:
                rd.pos += 34 * (n - i)
                break
            out.append(rd.g1())
        return n, out
    pk.A_len, pk.A_query = g1_vec()
    pk.B_domain = rd.dec()
    k = rd.dec()
    pk.B_indices = [rd.dec() for _ in range(k)]
    k2 = rd.dec()
    assert k2 == k
    pk.B_g2, pk.B_g1 = [], []
    for i in range(k):
        if max_points is not None and i >= max_points:
            rd.pos += (66 + 1 + 34 + 1) * (k - i)
            break
        pk.B_g2.append(rd.g2()); rd.expect(b" ")
        pk.B_g1.append(rd.g1()); rd.expect(b"\n")
    pk.H_len, pk.H_query = g1_vec()
    pk.L_len, pk.L_query = g1_vec()
    ni, na, nc = rd.dec(), rd.dec(), rd.dec()
    rows = []
    for _ in range(nc):
        row = []
        for _w in range(3):
            nt = rd.dec()
            row.append([(rd.dec(), rd.dec()) for _ in range(nt)])
        rows.append(tuple(row))
    pk.cs = R1CS(ni, na, rows)
    pk.end = rd.pos
    return pk


# ---------------------------------------------------------------------------------------------------------------
# the prover (r1cs_gg_ppzksnark.tcc:390-506)
def prover(pk, full, r, s):
    """full = primary || auxiliary (canonical ints); r, s canonical.  Returns (A, B, C) affine + the 5 MSM results."""
    cs = pk.cs
    n, ni = cs.num_variables, cs.num_inputs
    H = r1cs_to_qap_witness_map(cs, full)
    m = len(H) - 1
    cpa = [1] + list(full)
    jac1 = [G1.from_affine(a) for a in pk.A_query]
    eA = G1.multi_exp_with_mixed_addition(jac1[:n + 1], cpa[:n + 1])
    sc = [cpa[i] for i in pk.B_indices]
    eBg = G2.multi_exp_with_mixed_addition([G2.from_affine(a) for a in pk.B_g2], sc)
    eBh = G1.multi_exp_with_mixed_addition([G1.from_affine(a) for a in pk.B_g1], sc)
    eH = G1.multi_exp_inner([G1.from_affine(a) for a in pk.H_query[:m - 1]], H[:m - 1])
    eL = G1.multi_exp_with_mixed_addition([G1.from_affine(a) for a in pk.L_query], cpa[ni + 1:n + 1])
    d1, d2 = G1.from_affine(pk.delta_g1), G2.from_affine(pk.delta_g2)
    gA = G1.add(G1.add(G1.from_affine(pk.alpha_g1), eA), G1.mul(r, d1))
    g1B = G1.add(G1.add(G1.from_affine(pk.beta_g1), eBh), G1.mul(s, d1))
    g2B = G2.add(G2.add(G2.from_affine(pk.beta_g2), eBg), G2.mul(s, d2))
    gC = G1.add(G1.add(eH, eL), G1.add(G1.mul(s, gA), G1.mul(r, g1B)))
    gC = G1.add(gC, G1.neg(G1.mul(r * s % R_MOD, d1)))
    proof = (G1.to_affine(gA), G2.to_affine(g2B), G1.to_affine(gC))
    parts = dict(At=G1.to_affine(eA), Bt_g=G2.to_affine(eBg), Bt_h=G1.to_affine(eBh), Ht=G1.to_affine(eH), Lt=G1.to_affine(eL), H=H)
    return proof, parts


def proof_to_hex(proof):
    """mintcgo.cpp:112-187: A.x A.y B.x.c1 B.x.c0 B.y.c1 B.y.c0 C.x C.y, 64 lowercase hex chars each (big-endian).
    A point at infinity is printed as the affine form of (0,1,0), i.e. x=0,y=1 (alt_bn128_g1.cpp:60-77)."""
    A, B, C = proof
    A = A or (0, 1)
    C = C or (0, 1)
    B = B or (Fq2(0), Fq2(1))
    vals = [A[0], A[1], B[0].c1, B[0].c0, B[1].c1, B[1].c0, C[0], C[1]]
    return "".join("%064x" % v for v in vals)


# ---------------------------------------------------------------------------------------------------------------
# byte-level helpers of the cgo layer
def uint256S(s, nbytes=32):
    """base_blob::SetHex (uint256.h:200-226): skip 0x, read hex from the END into data[0..]; returns the in-memory bytes."""
    s = s.strip()
    if s[:2] in ("0x", "0X"):
        s = s[2:]
    digits = ""
    for ch in s:
        if ch in "0123456789abcdefABCDEF":
            digits += ch
        else:
            break
    out = bytearray(nbytes)
    i, pos = 0, len(digits) - 1
    while pos >= 0 and i < nbytes:
        out[i] = int(digits[pos], 16)
        pos -= 1
        if pos >= 0:
            out[i] |= int(digits[pos], 16) << 4
            pos -= 1
            i += 1
        else:
            break
    return bytes(out)


def blob_hex(b):
    """base_blob::GetHex: reversed bytes as hex (what ToString returns and the cgo helpers hand back)."""
    return bytes(reversed(b)).hex()


def arg_hex(b):
    """Encode in-memory blob bytes the way zktx.go passes them: '0x' + hex of the reversed bytes."""
    return "0x" + blob_hex(b)


def note_cm(value, sn, r):
    return hashlib.sha256(value.to_bytes(8, "little") + sn + r).digest()          # Note::cm (Note.h:30-44)


def notes_cm(value, pk20, r, sn_old):
    return hashlib.sha256(value.to_bytes(8, "little") + pk20 + r + sn_old).digest()  # NoteS::cm (Note.h:63-78)


def compute_prf(sk, r): return hashlib.sha256(sk + r).digest()                     # util.h:231-242
def compute_crh(pk20, r): return hashlib.sha256(pk20 + r).digest()                 # send/util.h:246-258


def sha256_compress(left, right):
    """SHA256Compress: one compression of the 64-byte block left||right from the standard IV, NO padding
    (deposit/IncrementalMerkleTree.cpp, CSHA256::FinalizeNoPadding)."""
    K = [0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
         0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
         0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
         0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
         0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
         0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
         0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]
    Hs = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
    M = 0xFFFFFFFF
    def rotr(x, n): return ((x >> n) | (x << (32 - n))) & M
    blk = left + right
    w = [int.from_bytes(blk[4 * i:4 * i + 4], "big") for i in range(16)]
    for i in range(16, 64):
        s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3)
        s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10)
        w.append((w[i - 16] + s0 + w[i - 7] + s1) & M)
    a, b, c, d, e, f, g, h = Hs
    for i in range(64):
        S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)
        ch = (e & f) ^ (~e & M & g)
        t1 = (h + S1 + ch + K[i] + w[i]) & M
        S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)
        mj = (a & b) ^ (a & c) ^ (b & c)
        t2 = (S0 + mj) & M
        h, g, f, e, d, c, b, a = g, f, e, (d + t1) & M, c, b, a, (t1 + t2) & M
    out = [(x + y) & M for x, y in zip(Hs, (a, b, c, d, e, f, g, h))]
    return b"".join(x.to_bytes(4, "big") for x in out)


def sha512_rng(idx):
    """libff::SHA512_rng<Fr> (libff/common/rng.tcc:26-72): SHA-512 of (idx:u64 LE || iter:u64 LE), the first 32 digest bytes read as a
    little-endian integer, the bits above the modulus' top bit (>= 254) cleared, next iter while the value is not below r.  The scalar
    stream of the synthetic kernel sweep; bases there are sha512_rng(2^32 + i) * generator."""
    import hashlib
    import struct
    it = 0
    while True:
        v = int.from_bytes(hashlib.sha512(struct.pack("<QQ", idx, it)).digest()[:32], "little") & ((1 << 254) - 1)
        if v < R_MOD:
            return v
        it += 1
