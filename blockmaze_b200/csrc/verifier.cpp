// verify{Mint,Send,Deposit,Redeem}proof of BlockMaze's cgo surface (SRC/*/*cgo.cpp verify*proof), i.e.
// r1cs_gg_ppzksnark_verifier_strong_IC (r1cs_gg_ppzksnark.tcc:524-623) over alt_bn128: public-input packing, a 4-6 term
// accumulation, three Miller loops and one final exponentiation.  ~10 ms in the reference and not on the prover's hot
// path, so it stays on the host; it must reproduce libff's reduced pairing EXACTLY because the verification key stores
// alpha_g1_beta_g2 as a GT element computed by libff:
//   Miller loop      alt_bn128_pairing.cpp:252-420   (flipped ate loop, loop count 6z+2, two Frobenius correction steps)
//   final exponent   alt_bn128_pairing.cpp:105-228   (easy part (q^6-1)(q^2+1); hard part a fixed multiple
//                                                      2z(6z^2+3z+1) of (q^4-q^2+1)/r, Fuentes-Castaneda et al.)
// Tower: Fq2 = Fq[u]/(u^2+1), Fq6 = Fq2[v]/(v^3 - xi), Fq12 = Fq6[w]/(w^2 - v), xi = 9 + u (alt_bn128_init.cpp:151-189).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include "../../include/zkb200.h"
#include "host_field.hpp"
#include "witness.hpp"

using namespace zkh;

namespace {

struct Fq6 {
    HFq2 c0, c1, c2;
    static Fq6 zero() { return Fq6{HFq2::zero(), HFq2::zero(), HFq2::zero()}; }
    static Fq6 one() { return Fq6{HFq2::one(), HFq2::zero(), HFq2::zero()}; }
    Fq6 operator+(const Fq6 &o) const { return Fq6{c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fq6 operator-(const Fq6 &o) const { return Fq6{c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fq6 neg() const { return Fq6{c0.neg(), c1.neg(), c2.neg()}; }
    bool operator==(const Fq6 &o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
};
static HFq2 mul_xi(const HFq2 &a) {          // (a0 + a1 u)(9 + u) = (9 a0 - a1) + (9 a1 + a0) u
    HFq a0_8 = a.c0.dbl().dbl().dbl(), a1_8 = a.c1.dbl().dbl().dbl();
    return HFq2{a0_8 + a.c0 - a.c1, a1_8 + a.c1 + a.c0};
}
static Fq6 mul(const Fq6 &a, const Fq6 &b) {
    const HFq2 v0 = a.c0 * b.c0, v1 = a.c1 * b.c1, v2 = a.c2 * b.c2;
    Fq6 r;
    r.c0 = v0 + mul_xi((a.c1 + a.c2) * (b.c1 + b.c2) - v1 - v2);
    r.c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - v0 - v1 + mul_xi(v2);
    r.c2 = (a.c0 + a.c2) * (b.c0 + b.c2) - v0 - v2 + v1;
    return r;
}
static Fq6 mul_by_v(const Fq6 &a) { return Fq6{mul_xi(a.c2), a.c0, a.c1}; }     // multiply by v: Fq12's non-residue
static Fq6 inverse(const Fq6 &a) {
    const HFq2 t0 = a.c0.sqr() - mul_xi(a.c1 * a.c2), t1 = mul_xi(a.c2.sqr()) - a.c0 * a.c1, t2 = a.c1.sqr() - a.c0 * a.c2;
    const HFq2 d = (a.c0 * t0 + mul_xi(a.c2 * t1 + a.c1 * t2)).inverse();
    return Fq6{t0 * d, t1 * d, t2 * d};
}
struct Fq12 {
    Fq6 c0, c1;
    static Fq12 one() { return Fq12{Fq6::one(), Fq6::zero()}; }
    bool operator==(const Fq12 &o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq12 conj() const { return Fq12{c0, c1.neg()}; }       // unitary_inverse
};
static Fq12 mul(const Fq12 &a, const Fq12 &b) {
    const Fq6 v0 = mul(a.c0, b.c0), v1 = mul(a.c1, b.c1);
    return Fq12{v0 + mul_by_v(v1), mul(a.c0 + a.c1, b.c0 + b.c1) - v0 - v1};
}
// complex squaring: (a0 + a1 w)^2 with w^2 = v  ->  2 Fq6 multiplications
static Fq12 sqr(const Fq12 &a) {
    const Fq6 ab = mul(a.c0, a.c1);
    return Fq12{mul(a.c0 + a.c1, a.c0 + mul_by_v(a.c1)) - ab - mul_by_v(ab), ab + ab};
}
// Squaring in the cyclotomic subgroup (every element after the easy part of the final exponentiation): Granger-Scott over
// Fq4 = Fq2[y]/(y^2 - xi), y = w^3.  alpha = A + B w + C w^2 with A = (c0.c0, c1.c1), B = (c1.c0, c0.c2), C = (c0.c1, c1.c2);
// alpha^2 = (3 A^2 - 2 conj A) + (3 y C^2 + 2 conj B) w + (3 B^2 - 2 conj C) w^2.  Three Fq4 squarings = 9 Fq2 multiplications.
static void fq4_sqr(const HFq2 &a, const HFq2 &b, HFq2 &even, HFq2 &odd) {
    const HFq2 ab = a * b;
    even = (a + b) * (a + mul_xi(b)) - ab - mul_xi(ab);
    odd = ab + ab;
}
static HFq2 three_minus_two(const HFq2 &t, const HFq2 &z) { const HFq2 d = t - z; return d + d + t; }     // 3t - 2z
static HFq2 three_plus_two(const HFq2 &t, const HFq2 &z) { const HFq2 d = t + z; return d + d + t; }      // 3t + 2z
static Fq12 cyclotomic_sqr(const Fq12 &a) {
    HFq2 a0, a1, b0, b1, c0, c1;
    fq4_sqr(a.c0.c0, a.c1.c1, a0, a1);          // A^2
    fq4_sqr(a.c1.c0, a.c0.c2, b0, b1);          // B^2
    fq4_sqr(a.c0.c1, a.c1.c2, c0, c1);          // C^2
    Fq12 r;
    r.c0.c0 = three_minus_two(a0, a.c0.c0); r.c1.c1 = three_plus_two(a1, a.c1.c1);
    r.c1.c0 = three_plus_two(mul_xi(c1), a.c1.c0); r.c0.c2 = three_minus_two(c0, a.c0.c2);
    r.c0.c1 = three_minus_two(b0, a.c0.c1); r.c1.c2 = three_plus_two(b1, a.c1.c2);
    return r;
}
// f * (x0 + x2 v^2 + x4 v w): the sparse line element of the Miller loop (libff's mul_by_024), 15 Fq2 multiplications
static Fq6 mul_by_02(const Fq6 &a, const HFq2 &x0, const HFq2 &x2) {
    return Fq6{a.c0 * x0 + mul_xi(a.c1 * x2), a.c1 * x0 + mul_xi(a.c2 * x2), a.c2 * x0 + a.c0 * x2};
}
static Fq6 mul_by_1(const Fq6 &a, const HFq2 &x1) { return Fq6{mul_xi(a.c2 * x1), a.c0 * x1, a.c1 * x1}; }
static Fq12 mul_sparse(const Fq12 &f, const HFq2 &x0, const HFq2 &x2, const HFq2 &x4) {
    const Fq6 v0 = mul_by_02(f.c0, x0, x2), v1 = mul_by_1(f.c1, x4);
    return Fq12{v0 + mul_by_v(v1), mul(f.c0 + f.c1, Fq6{x0, x4, x2}) - v0 - v1};
}
static Fq12 inverse(const Fq12 &a) {
    const Fq6 d = inverse(mul(a.c0, a.c0) - mul_by_v(mul(a.c1, a.c1)));
    return Fq12{mul(a.c0, d), mul(a.c1, d).neg()};
}

// Frobenius constants, derived at start-up from gamma = xi^((q-1)/6) instead of being tabulated (alt_bn128_init.cpp:156-189)
struct Frob {
    HFq2 g1, g2, g3, g4, g5;      // gamma^k
    Frob() {
        uint64_t e[4]; memcpy(e, FqTag::MOD, 32); e[0] -= 1;          // q - 1
        unsigned __int128 rem = 0;                                      // divide by 6
        for (int i = 3; i >= 0; i--) { unsigned __int128 cur = (rem << 64) | e[i]; e[i] = (uint64_t)(cur / 6); rem = cur % 6; }
        const HFq2 xi{HFq::from_u64(9), HFq::one()};
        HFq2 r = HFq2::one();
        for (int i = 255; i >= 0; i--) { r = r.sqr(); if ((e[i >> 6] >> (i & 63)) & 1) r = r * xi; }
        g1 = r; g2 = g1 * g1; g3 = g2 * g1; g4 = g2 * g2; g5 = g4 * g1;
    }
};
static const Frob &frob() { static Frob f; return f; }
static HFq2 conj2(const HFq2 &a) { return HFq2{a.c0, a.c1.neg()}; }
// x -> x^q on Fq12 (Fp12_2over3over2_model::Frobenius_map(1), fp12_2over3over2.tcc:159-163)
static Fq12 frobenius(const Fq12 &a) {
    const Frob &F = frob();
    Fq12 r;
    r.c0 = Fq6{conj2(a.c0.c0), conj2(a.c0.c1) * F.g2, conj2(a.c0.c2) * F.g4};
    r.c1 = Fq6{conj2(a.c1.c0) * F.g1, conj2(a.c1.c1) * F.g3, conj2(a.c1.c2) * F.g5};
    return r;
}
static Fq12 pow_z(const Fq12 &a) {           // a^z, z = 4965661367192848881 (alt_bn128_init.cpp:327)
    const uint64_t z = 4965661367192848881ull;
    Fq12 r = Fq12::one();
    for (int i = 63; i >= 0; i--) { r = cyclotomic_sqr(r); if ((z >> i) & 1) r = mul(r, a); }       // a is in the cyclotomic subgroup
    return r;
}
static Fq12 final_exponentiation(const Fq12 &elt) {
    // first chunk: elt^((q^6-1)(q^2+1))
    const Fq12 C = mul(elt.conj(), inverse(elt));
    const Fq12 f = mul(frobenius(frobenius(C)), C);
    // last chunk (alt_bn128_pairing.cpp:137-228); exp_by_neg_z = conj(pow_z) on the cyclotomic subgroup
    const Fq12 A = pow_z(f).conj(), B = cyclotomic_sqr(A), Cc = cyclotomic_sqr(B), D = mul(Cc, B), E = pow_z(D).conj(), Fv = cyclotomic_sqr(E),
               G = pow_z(Fv).conj();
    const Fq12 H = D.conj(), I = G.conj(), J = mul(I, E), K = mul(J, H), L = mul(K, B), M = mul(K, E), N = mul(M, f);
    const Fq12 O = frobenius(L), P = mul(O, N), Q = frobenius(frobenius(K)), R = mul(Q, P), S = f.conj(), T = mul(S, L);
    const Fq12 U = frobenius(frobenius(frobenius(T)));
    return mul(U, R);
}

struct EllCoeffs { HFq2 ell_0, ell_VW, ell_VV; };
struct G2Proj { HFq2 X, Y, Z; };
static HFq2 twist_b() { static HFq2 b = HFq2{HFq::from_u64(3), HFq::zero()} * HFq2{HFq::from_u64(9), HFq::one()}.inverse(); return b; }
static HFq2 scale(const HFq2 &a, const HFq &s) { return HFq2{a.c0 * s, a.c1 * s}; }
static void doubling_step(G2Proj &cur, EllCoeffs &c) {           // alt_bn128_pairing.cpp:252-278
    static const HFq two_inv = HFq::from_u64(2).inverse();
    static const HFq2 xi{HFq::from_u64(9), HFq::one()};
    const HFq2 X = cur.X, Y = cur.Y, Z = cur.Z;
    const HFq2 A = scale(X * Y, two_inv), B = Y.sqr(), C = Z.sqr(), D = C + C + C, E = twist_b() * D, F = E + E + E;
    const HFq2 G = scale(B + F, two_inv), H = (Y + Z).sqr() - (B + C), I = E - B, J = X.sqr(), E2 = E.sqr();
    cur.X = A * (B - F); cur.Y = G.sqr() - (E2 + E2 + E2); cur.Z = B * H;
    c.ell_0 = xi * I; c.ell_VW = H.neg(); c.ell_VV = J + J + J;
}
static void addition_step(const HFq2 &x2, const HFq2 &y2, G2Proj &cur, EllCoeffs &c) {      // alt_bn128_pairing.cpp:280-302
    static const HFq2 xi{HFq::from_u64(9), HFq::one()};
    const HFq2 X1 = cur.X, Y1 = cur.Y, Z1 = cur.Z;
    const HFq2 D = X1 - x2 * Z1, E = Y1 - y2 * Z1, F = D.sqr(), G = E.sqr(), H = D * F, I = X1 * F, J = H + Z1 * G - (I + I);
    cur.X = D * J; cur.Y = E * (I - J) - (H * Y1); cur.Z = Z1 * H;
    c.ell_0 = xi * (E * x2 - D * y2); c.ell_VV = E.neg(); c.ell_VW = D;
}
static Fq12 line_mul(const Fq12 &f, const EllCoeffs &c, const HFq &px, const HFq &py) {
    // mul_by_024(ell_0, PY*ell_VW, PX*ell_VV): the sparse factor is (c0 = (ell_0, 0, ell_VV'), c1 = (0, ell_VW', 0))  (fp12_2over3over2.tcc:244-248)
    return mul_sparse(f, c.ell_0, scale(c.ell_VV, px), scale(c.ell_VW, py));
}
// Line coefficients of the whole Miller loop for one G2 point (libff's alt_bn128_ate_precompute_G2): they depend on Q only, so the two
// G2 points of a verification key (gamma, delta) are prepared once per key and only the proof's B is prepared per verification.
static const unsigned __int128 ATE_LOOP = ((unsigned __int128)1 << 64) | 0x9d797039be763ba8ull;   // 29793968203157093288 = 6z+2 (alt_bn128_init.cpp:324)
struct G2Prepared { std::vector<EllCoeffs> coeffs; };
static G2Prepared prepare_g2(const HG2Affine &Q) {
    const Frob &F = frob();
    G2Prepared out;
    out.coeffs.reserve(104);
    G2Proj R{Q.x, Q.y, HFq2::one()};
    EllCoeffs c;
    for (int i = 63; i >= 0; i--) {            // bit 64 is the MSB and is skipped
        doubling_step(R, c); out.coeffs.push_back(c);
        if ((ATE_LOOP >> i) & 1) { addition_step(Q.x, Q.y, R, c); out.coeffs.push_back(c); }
    }
    // Q1 = pi(Q), Q2 = -pi^2(Q)  (alt_bn128_G2::mul_by_q: x^q * xi^((q-1)/3), y^q * xi^((q-1)/2); alt_bn128_g2.cpp)
    const HFq2 q1x = conj2(Q.x) * F.g2, q1y = conj2(Q.y) * F.g3;
    const HFq2 q2x = conj2(q1x) * F.g2, q2y = (conj2(q1y) * F.g3).neg();
    addition_step(q1x, q1y, R, c); out.coeffs.push_back(c);
    addition_step(q2x, q2y, R, c); out.coeffs.push_back(c);
    return out;
}
// product of Miller loops over (P_i, prepared Q_i), affine P, none at infinity: one Fq12 squaring per bit for all pairs
static Fq12 multi_miller(const std::vector<HG1Affine> &Ps, const std::vector<const G2Prepared *> &Qs) {
    const size_t n = Ps.size();
    Fq12 f = Fq12::one();
    size_t idx = 0;
    for (int i = 63; i >= 0; i--) {
        f = sqr(f);
        for (size_t k = 0; k < n; k++) f = line_mul(f, Qs[k]->coeffs[idx], Ps[k].x, Ps[k].y);
        idx++;
        if ((ATE_LOOP >> i) & 1) {
            for (size_t k = 0; k < n; k++) f = line_mul(f, Qs[k]->coeffs[idx], Ps[k].x, Ps[k].y);
            idx++;
        }
    }
    for (int e = 0; e < 2; e++, idx++)
        for (size_t k = 0; k < n; k++) f = line_mul(f, Qs[k]->coeffs[idx], Ps[k].x, Ps[k].y);
    return f;
}
static Fq12 multi_miller(const std::vector<HG1Affine> &Ps, const std::vector<HG2Affine> &Qs) {
    std::vector<G2Prepared> prep; prep.reserve(Qs.size());
    std::vector<const G2Prepared *> ptr;
    for (const HG2Affine &q : Qs) { prep.push_back(prepare_g2(q)); ptr.push_back(&prep.back()); }
    return multi_miller(Ps, ptr);
}

// ---- verification key ---------------------------------------------------------------------------------------------------
static bool fq_sqrt(const HFq &a, HFq &out) {
    uint64_t e[4]; memcpy(e, FqTag::MOD, 32); e[0] += 1;
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 2) | (i < 3 ? e[i + 1] << 62 : 0);
    out = a.pow(e);
    return out.sqr() == a;
}
static bool fq2_sqrt(const HFq2 &a, HFq2 &out) {
    HFq t;
    if (a.c1.is_zero()) {
        if (fq_sqrt(a.c0, t)) { out = HFq2{t, HFq::zero()}; return true; }
        if (fq_sqrt(a.c0.neg(), t)) { out = HFq2{HFq::zero(), t}; return true; }
        return false;
    }
    HFq n;
    if (!fq_sqrt(a.c0.sqr() + a.c1.sqr(), n)) return false;
    const HFq half = HFq::from_u64(2).inverse();
    HFq x0;
    if (!fq_sqrt((a.c0 + n) * half, x0) && !fq_sqrt((a.c0 - n) * half, x0)) return false;
    out = HFq2{x0, a.c1 * x0.dbl().inverse()};
    return out.sqr() == a;
}
static bool lsb(const HFq &x) { uint64_t c[4]; x.to_canonical(c); return c[0] & 1; }

// Fixed-base table of one gamma_ABC point: tab[k][d-1] = d * 16^k * P (k < 64, d = 1..15), so input * P is 64 additions, no doublings
struct FixedBase { std::vector<HG1> tab; };
static FixedBase fixed_base(const HG1Affine &P) {
    FixedBase f; f.tab.resize(64 * 15);
    HG1 base = HG1::from_affine(P);
    for (int k = 0; k < 64; k++) {
        HG1 cur = base;
        for (int d = 1; d <= 15; d++) { f.tab[k * 15 + d - 1] = cur; cur = cur.add(base); }
        base = cur;                                         // 16 * base
    }
    return f;
}
struct VerificationKey {
    Fq12 alpha_beta;
    HG2Affine gamma_g2, delta_g2;
    std::vector<HG1Affine> gamma_abc;        // [0] = first, then one per public input
    // derived once per key (libsnark's "processed" verification key goes half of this way)
    G2Prepared gamma_prep, delta_prep;
    std::vector<FixedBase> abc_tab;          // for gamma_abc[1..]
    bool ok = false;
};
struct VkReader {
    const std::string &d; size_t p = 0; bool fail = false;
    explicit VkReader(const std::string &s) : d(s) {}
    HFq dec_fq() {
        size_t b = p;
        while (p < d.size() && d[p] >= '0' && d[p] <= '9') p++;
        HFq v = HFq::zero();
        if (p == b || !HFq::from_dec(d.data() + b, p - b, v)) fail = true;
        if (p < d.size()) p++;           // separator
        return v;
    }
    uint64_t dec_u64() {
        uint64_t v = 0; size_t b = p;
        while (p < d.size() && d[p] >= '0' && d[p] <= '9') v = v * 10 + (uint64_t)(d[p++] - '0');
        if (p == b) fail = true;
        if (p < d.size()) p++;
        return v;
    }
    void expect(char c) { if (p < d.size() && d[p] == c) p++; else fail = true; }
    HG1Affine g1() {          // alt_bn128_g1.cpp:420-465
        if (p + 34 > d.size()) { fail = true; return HG1Affine::inf(); }
        const bool inf = d[p] == '1', ybit = d[p + 33] == '1';
        HFq x; memcpy(x.v, d.data() + p + 1, 32); p += 34;
        if (inf) return HG1Affine::inf();
        HFq y;
        if (!fq_sqrt(x.sqr() * x + HFq::from_u64(3), y)) { fail = true; return HG1Affine::inf(); }
        if (lsb(y) != ybit) y = y.neg();
        return HG1Affine{x, y};
    }
    HG2Affine g2() {          // alt_bn128_g2.cpp:433-478
        if (p + 66 > d.size()) { fail = true; return HG2Affine::inf(); }
        const bool inf = d[p] == '1', ybit = d[p + 65] == '1';
        HFq2 x; memcpy(x.c0.v, d.data() + p + 1, 32); memcpy(x.c1.v, d.data() + p + 33, 32); p += 66;
        if (inf) return HG2Affine::inf();
        HFq2 y;
        if (!fq2_sqrt(x.sqr() * x + twist_b(), y)) { fail = true; return HG2Affine::inf(); }
        if (lsb(y.c0) != ybit) y = y.neg();
        return HG2Affine{x, y};
    }
};
// vk grammar (r1cs_gg_ppzksnark.tcc:99-108, accumulation_vector.tcc:63-69, sparse_vector.tcc:272-288)
static VerificationKey parse_vk(const std::string &data) {
    VerificationKey vk;
    VkReader r(data);
    HFq2 *slots[6] = {&vk.alpha_beta.c0.c0, &vk.alpha_beta.c0.c1, &vk.alpha_beta.c0.c2, &vk.alpha_beta.c1.c0, &vk.alpha_beta.c1.c1, &vk.alpha_beta.c1.c2};
    for (auto *s : slots) { s->c0 = r.dec_fq(); s->c1 = r.dec_fq(); }
    vk.gamma_g2 = r.g2(); r.expect('\n');
    vk.delta_g2 = r.g2(); r.expect('\n');
    vk.gamma_abc.push_back(r.g1()); r.expect('\n');
    r.dec_u64();                                   // domain size
    const uint64_t k = r.dec_u64();
    if (k > 64) r.fail = true;
    for (uint64_t i = 0; i < k && !r.fail; i++) r.dec_u64();
    const uint64_t k2 = r.dec_u64();
    if (k2 != k) r.fail = true;
    for (uint64_t i = 0; i < k && !r.fail; i++) { vk.gamma_abc.push_back(r.g1()); r.expect('\n'); }
    vk.ok = !r.fail;
    if (vk.ok) {
        vk.gamma_prep = prepare_g2(vk.gamma_g2); vk.delta_prep = prepare_g2(vk.delta_g2);
        for (size_t i = 1; i < vk.gamma_abc.size(); i++) vk.abc_tab.push_back(fixed_base(vk.gamma_abc[i]));
    }
    return vk;
}

static std::mutex g_vk_mu;
static std::shared_ptr<const VerificationKey> g_vk[4];
static std::string g_vk_dir[4];
static const char *NAMES[4] = {"mint", "send", "deposit", "redeem"};
static std::string key_dir_now() { return zkw::key_dir(); }

static bool hex_fq(const char *s, HFq &out) {      // 64 lowercase hex chars, big-endian (libsnarkBigintFromBytes, mintcgo.cpp:36-48)
    uint64_t v[4] = {0, 0, 0, 0};
    for (int i = 0; i < 64; i++) {
        const char ch = s[i];
        int d;
        if (ch >= '0' && ch <= '9') d = ch - '0'; else if (ch >= 'a' && ch <= 'f') d = ch - 'a' + 10; else return false;
        v[3 - i / 16] |= (uint64_t)d << (4 * (15 - i % 16));
    }
    // The reference builds Fq from the raw 256-bit integer (libsnarkBigintFromBytes -> Fp_model(bigint), fp.tcc:190-194: mul_reduce by
    // R^2), i.e. it REDUCES mod q: an encoding such as x + q is accepted by reference nodes, so it must be accepted here as well or a
    // mixed network forks.  R^2 is the left operand so that the row accumulator stays below 2q for any 256-bit v.
    out = HFq::raw(FqTag::R2) * HFq::raw(v);
    return true;
}
// pack_bit_vector_into_field_element_vector (field_utils.tcc:79-103): chunks of 253 bits, bit j of a chunk -> 2^j
static std::vector<HFr> pack_bits(const std::vector<bool> &bits) {
    std::vector<HFr> out;
    for (size_t c = 0; c * 253 < bits.size(); c++) {
        uint64_t x[4] = {0, 0, 0, 0};
        for (size_t k = 0; k < 253 && c * 253 + k < bits.size(); k++) if (bits[c * 253 + k]) x[k >> 6] |= (uint64_t)1 << (k & 63);
        out.push_back(HFr::from_canonical(x));
    }
    return out;
}
static void push_blob(std::vector<bool> &b, const uint8_t *blob, size_t nbytes) {
    for (size_t i = 0; i < nbytes * 8; i++) b.push_back((blob[i >> 3] >> (7 - (i & 7))) & 1);
}
static void push_u64(std::vector<bool> &b, uint64_t v) { uint8_t le[8]; for (int i = 0; i < 8; i++) le[i] = (uint8_t)(v >> (8 * i)); push_blob(b, le, 8); }

static std::shared_ptr<const VerificationKey> vk_for(int circuit) {        // concurrent verifications share the key and do not hold the lock
    std::lock_guard<std::mutex> lk(g_vk_mu);
    const std::string dir = key_dir_now();
    if (!g_vk[circuit] || !g_vk[circuit]->ok || g_vk_dir[circuit] != dir) {
        std::ifstream fh(dir + "/" + NAMES[circuit] + "vk.txt", std::ios::binary);
        std::string data((std::istreambuf_iterator<char>(fh)), std::istreambuf_iterator<char>());
        g_vk[circuit] = std::make_shared<const VerificationKey>(parse_vk(data));
        g_vk_dir[circuit] = dir;
        if (!g_vk[circuit]->ok) { fprintf(stderr, "zkb200: cannot read verification key %s/%svk.txt\n", dir.c_str(), NAMES[circuit]); return nullptr; }
    }
    return g_vk[circuit];
}
struct ParsedProof { HG1Affine A, C; HG2Affine B; };
// proof string -> points; false when the string is malformed or a point is off its curve (proof.is_well_formed(), r1cs_gg_ppzksnark.tcc:537-541)
static bool parse_proof(const char *proof, ParsedProof &out) {
    if (!proof || strnlen(proof, 512) < 512) return false;
    HFq c[8];
    for (int i = 0; i < 8; i++) if (!hex_fq(proof + 64 * i, c[i])) return false;
    out.A = HG1Affine{c[0], c[1]}; out.C = HG1Affine{c[6], c[7]};
    out.B = HG2Affine{HFq2{c[3], c[2]}, HFq2{c[5], c[4]}};            // string order is c1 then c0 (mintcgo.cpp:150-169)
    const HFq three = HFq::from_u64(3);
    if (!(out.A.y.sqr() == out.A.x.sqr() * out.A.x + three) || !(out.C.y.sqr() == out.C.x.sqr() * out.C.x + three)) return false;
    return out.B.y.sqr() == out.B.x.sqr() * out.B.x + twist_b();
}
// sum of k[i] * gamma_ABC[i + 1] over the fixed-base tables of the key (accumulate(), r1cs_gg_ppzksnark.tcc:549-553)
static HG1 abc_combination(const VerificationKey &vk, const std::vector<HFr> &k_mont, HG1 acc) {
    for (size_t i = 0; i < k_mont.size(); i++) {
        uint64_t k[4]; k_mont[i].to_canonical(k);
        const FixedBase &fb = vk.abc_tab[i];
        for (int w = 0; w < 64; w++) {
            const unsigned d = (unsigned)((k[w >> 4] >> ((w & 15) * 4)) & 15);
            if (d) acc = acc.add(fb.tab[w * 15 + d - 1]);
        }
    }
    return acc;
}
static bool check_one(const VerificationKey &vk, const ParsedProof &pp, const std::vector<HFr> &inputs) {
    if (inputs.size() + 1 != vk.gamma_abc.size()) return false;
    const HG1Affine acc_a = abc_combination(vk, inputs, HG1::from_affine(vk.gamma_abc[0])).to_affine();
    // e(A, B) == alpha_beta * e(acc, gamma) * e(C, delta)   <=>   FE( ML(A,B) * ML(-acc,gamma) * ML(-C,delta) ) == alpha_beta:
    // one Miller loop over the three pairs (one Fq12 squaring per bit for all of them), one final exponentiation
    const G2Prepared b_prep = prepare_g2(pp.B);
    std::vector<HG1Affine> Ps{pp.A}; std::vector<const G2Prepared *> Qs{&b_prep};
    if (!acc_a.is_inf()) { Ps.push_back(HG1Affine{acc_a.x, acc_a.y.neg()}); Qs.push_back(&vk.gamma_prep); }
    Ps.push_back(HG1Affine{pp.C.x, pp.C.y.neg()}); Qs.push_back(&vk.delta_prep);
    return final_exponentiation(multi_miller(Ps, Qs)) == vk.alpha_beta;
}
static bool verify(int circuit, const char *proof, const std::vector<bool> &input_bits) {
    const std::shared_ptr<const VerificationKey> held = vk_for(circuit);
    if (!held) return false;
    ParsedProof pp;
    if (!parse_proof(proof, pp)) return false;
    return check_one(*held, pp, pack_bits(input_bits));
}

// public-input bits of one verify<Circuit>proof call: the blobs in argument order (value_s last for mint / redeem), exactly as the
// four entry points below pack them
static std::vector<bool> input_bits_of(int circuit, const char *const *s, uint64_t value_s) {
    std::vector<bool> bits;
    uint8_t blob[32];
    if (circuit == ZKB200_DEPOSIT) {            // RT, pk (20 bytes), cmtB_old, sn_old, cmtB, sn_s
        for (int i = 0; i < 6; i++) { const size_t nb = i == 1 ? 20 : 32; zkw::parse_hex_blob(s[i], blob, nb); push_blob(bits, blob, nb); }
        return bits;
    }
    const int n = circuit == ZKB200_SEND ? 4 : 3;
    for (int i = 0; i < n; i++) { zkw::parse_hex_blob(s[i], blob, 32); push_blob(bits, blob, 32); }
    if (circuit != ZKB200_SEND) push_u64(bits, value_s);
    return bits;
}

// ---- batch verification ------------------------------------------------------------------------------------------------------------
// Every BlockMaze node verifies each zk transaction twice, in the pool and at block import (core/tx_pool.go:619-641,
// core/state_processor.go:113-158), one r1cs_gg_ppzksnark_verifier_strong_IC call each (r1cs_gg_ppzksnark.tcc:524-623): three Miller loops and
// a final exponentiation per proof.  A block's worth of proofs is checked here with ONE final exponentiation: with random 128-bit z_i,
//     prod_i e(z_i A_i, B_i) * prod_c [ e(-sum_i z_i acc_i, gamma_c) * e(-sum_i z_i C_i, delta_c) ]  ==  prod_c alpha_beta_c ^ (sum_i z_i)
// (c runs over the circuits present; sum_i z_i acc_i is one fixed-base combination per circuit because the z_i fold into the input
// scalars).  Per proof that leaves one Miller loop and two 128-bit G1 multiplications.  The verdicts are those of the per-proof check:
// malformed proofs are rejected up front as there, a B outside the order-r subgroup (where the pairing is not bilinear and the
// combination proves nothing) is sent to the per-proof check, and if the combined equation fails every proof is checked on its own, so
// the only difference is a 2^-128 chance of accepting a batch that holds an invalid proof.
// Membership of Q in E'(Fq2) in the order-r subgroup.  By definition [r]Q = O; on a BN curve the twisted Frobenius psi acts on G2 as
// multiplication by q = t - 1 = 6z^2 (mod r), and psi(Q) = [6z^2]Q holds ONLY on G2 (El Housni, Guillevic, Piellard, "Co-factor clearing and
// subgroup membership testing on pairing-friendly curves", 2022, section 4.3), so a 127-bit multiplication replaces the 254-bit one.
// psi = alt_bn128_G2::mul_by_q (alt_bn128_g2.cpp: x^q * xi^((q-1)/3), y^q * xi^((q-1)/2)), the map prepare_g2() uses for Q1.
static const uint64_t FR_ORDER[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
static const uint64_t SIX_Z_SQUARED[4] = {0xf83e9682e87cfd46ull, 0x6f4d8248eeb859fbull, 0, 0};       // 6 * 4965661367192848881^2
static bool g2_in_subgroup_by_order(const HG2Affine &Q) { return HG2::from_affine(Q).mul(FR_ORDER).is_inf(); }
static bool g2_in_subgroup(const HG2Affine &Q) {
    const Frob &F = frob();
    const HG2Affine psi{conj2(Q.x) * F.g2, conj2(Q.y) * F.g3};
    const HG2Affine m = HG2::from_affine(Q).mul(SIX_Z_SQUARED).to_affine();
    return !m.is_inf() && m.x == psi.x && m.y == psi.y;
}
static Fq12 pow_cyclotomic(const Fq12 &a, const uint64_t e[4]) {          // a in the cyclotomic subgroup (a pairing value)
    Fq12 r = Fq12::one();
    bool started = false;
    for (int i = 255; i >= 0; i--) {
        if (started) r = cyclotomic_sqr(r);
        if ((e[i >> 6] >> (i & 63)) & 1) { r = started ? mul(r, a) : a; started = true; }
    }
    return r;
}
struct BatchItem { int circuit; ParsedProof pp; std::vector<HFr> inputs; uint64_t z[2]; };
struct BatchPartial {                       // what one worker thread contributes
    Fq12 f = Fq12::one();
    HG1 sumC[4] = {HG1::inf(), HG1::inf(), HG1::inf(), HG1::inf()};
};
static void batch_worker(const std::vector<BatchItem *> &items, size_t lo, size_t hi, BatchPartial &out) {
    // Miller loops of a few proofs at a time share their Fq12 squarings
    const size_t GROUP = 8;
    for (size_t g = lo; g < hi; g += GROUP) {
        const size_t ge = g + GROUP < hi ? g + GROUP : hi;
        std::vector<G2Prepared> prep; prep.reserve(ge - g);
        std::vector<HG1Affine> Ps; std::vector<const G2Prepared *> Qs;
        for (size_t i = g; i < ge; i++) {
            const BatchItem &it = *items[i];
            const uint64_t z[4] = {it.z[0], it.z[1], 0, 0};
            Ps.push_back(HG1::from_affine(it.pp.A).mul(z).to_affine());
            prep.push_back(prepare_g2(it.pp.B));
            out.sumC[it.circuit] = out.sumC[it.circuit].add(HG1::from_affine(it.pp.C).mul(z));
        }
        for (const G2Prepared &p : prep) Qs.push_back(&p);
        std::vector<HG1Affine> P2; std::vector<const G2Prepared *> Q2;
        for (size_t k = 0; k < Ps.size(); k++) if (!Ps[k].is_inf()) { P2.push_back(Ps[k]); Q2.push_back(Qs[k]); }
        if (!P2.empty()) out.f = mul(out.f, multi_miller(P2, Q2));
    }
}
} // namespace

// test hook: both membership tests on one point (128 B canonical, on the twist curve); bit 0 = endomorphism test, bit 1 = [r]Q = O
int zkb200_g2_subgroup_check(const uint8_t point[128]) {
    uint64_t c[16]; memcpy(c, point, 128);
    const HG2Affine Q{HFq2{HFq::from_canonical(c), HFq::from_canonical(c + 4)}, HFq2{HFq::from_canonical(c + 8), HFq::from_canonical(c + 12)}};
    return (g2_in_subgroup(Q) ? 1 : 0) | (g2_in_subgroup_by_order(Q) ? 2 : 0);
}

int zkb200_verify_batch(size_t n, const zkb200_vtx *items, uint8_t *ok, int threads) {
    if (n && (!items || !ok)) return -1;
    for (size_t i = 0; i < n; i++) { ok[i] = 0; if (items[i].circuit < 0 || items[i].circuit > 3) return -1; }
    std::shared_ptr<const VerificationKey> vks[4];
    std::vector<BatchItem> parsed(n);
    std::vector<BatchItem *> batch;                 // well-formed, B in the subgroup, input count right
    std::vector<size_t> batch_idx, single_idx;
    std::random_device rd;
    for (size_t i = 0; i < n; i++) {
        const int c = items[i].circuit;
        if (!vks[c]) vks[c] = vk_for(c);
        if (!vks[c]) return -1;
        BatchItem &it = parsed[i];
        it.circuit = c;
        if (!parse_proof(items[i].proof, it.pp)) continue;                          // verdict 0, as the per-proof check
        it.inputs = pack_bits(input_bits_of(c, items[i].s, items[i].value_s));
        if (it.inputs.size() + 1 != vks[c]->gamma_abc.size()) continue;
        if (!g2_in_subgroup(it.pp.B)) { single_idx.push_back(i); continue; }
        do { it.z[0] = ((uint64_t)rd() << 32) | rd(); it.z[1] = ((uint64_t)rd() << 32) | rd(); } while (!(it.z[0] | it.z[1]));
        batch.push_back(&it); batch_idx.push_back(i);
    }
    bool batch_ok = true;
    if (batch.size() == 1) { single_idx.push_back(batch_idx[0]); batch.clear(); batch_idx.clear(); }
    if (!batch.empty()) {
        if (threads <= 0) { threads = (int)std::thread::hardware_concurrency(); if (threads > 32) threads = 32; if (threads < 1) threads = 1; }
        if ((size_t)threads > (batch.size() + 3) / 4) threads = (int)((batch.size() + 3) / 4);
        std::vector<BatchPartial> part(threads);
        std::vector<std::thread> pool;
        const size_t per = (batch.size() + threads - 1) / threads;
        for (int t = 1; t < threads; t++) {
            const size_t lo = t * per, hi = lo + per < batch.size() ? lo + per : batch.size();
            if (lo < hi) pool.emplace_back(batch_worker, std::cref(batch), lo, hi, std::ref(part[t]));
        }
        batch_worker(batch, 0, per < batch.size() ? per : batch.size(), part[0]);
        for (auto &th : pool) th.join();
        Fq12 f = Fq12::one();
        HG1 sumC[4] = {HG1::inf(), HG1::inf(), HG1::inf(), HG1::inf()};
        for (const BatchPartial &p : part) { f = mul(f, p.f); for (int c = 0; c < 4; c++) sumC[c] = sumC[c].add(p.sumC[c]); }
        // per circuit: the combined scalars of gamma_ABC (Fr arithmetic), then the two pairs against the prepared gamma and delta
        Fq12 rhs = Fq12::one();
        std::vector<HG1Affine> Ps; std::vector<const G2Prepared *> Qs;
        for (int c = 0; c < 4; c++) {
            bool any = false;
            for (const BatchItem *it : batch) if (it->circuit == c) { any = true; break; }
            if (!any) continue;
            const VerificationKey &vk = *vks[c];
            HFr zsum = HFr::zero();
            std::vector<HFr> comb(vk.gamma_abc.size() - 1, HFr::zero());
            for (const BatchItem *it : batch) if (it->circuit == c) {
                const uint64_t z4[4] = {it->z[0], it->z[1], 0, 0};
                const HFr z = HFr::from_canonical(z4);
                zsum = zsum + z;
                for (size_t j = 0; j < comb.size(); j++) comb[j] = comb[j] + z * it->inputs[j];
            }
            uint64_t zs[4]; zsum.to_canonical(zs);
            const HG1Affine acc = abc_combination(vk, comb, HG1::from_affine(vk.gamma_abc[0]).mul(zs)).to_affine();
            if (!acc.is_inf()) { Ps.push_back(HG1Affine{acc.x, acc.y.neg()}); Qs.push_back(&vk.gamma_prep); }
            const HG1Affine sc = sumC[c].to_affine();
            if (!sc.is_inf()) { Ps.push_back(HG1Affine{sc.x, sc.y.neg()}); Qs.push_back(&vk.delta_prep); }
            rhs = mul(rhs, pow_cyclotomic(vk.alpha_beta, zs));
        }
        if (!Ps.empty()) f = mul(f, multi_miller(Ps, Qs));
        batch_ok = final_exponentiation(f) == rhs;
    }
    if (batch_ok) for (size_t i : batch_idx) ok[i] = 1;
    else single_idx.insert(single_idx.end(), batch_idx.begin(), batch_idx.end());      // at least one is bad: find out which
    // per-proof checks (the reference's verdict by construction), spread over the same number of threads
    if (!single_idx.empty()) {
        if (threads <= 0) threads = 1;
        const int nt = (size_t)threads < single_idx.size() ? threads : (int)single_idx.size();
        std::vector<std::thread> pool;
        auto work = [&](int t) { for (size_t k = t; k < single_idx.size(); k += nt) { const size_t i = single_idx[k]; ok[i] = check_one(*vks[parsed[i].circuit], parsed[i].pp, parsed[i].inputs) ? 1 : 0; } };
        for (int t = 1; t < nt; t++) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
    }
    int good = 0;
    for (size_t i = 0; i < n; i++) good += ok[i];
    return good;
}

namespace {
} // namespace

// e(P, Q) as libff's alt_bn128_reduced_pairing computes it; out = the 12 Fq coefficients in the order operator<< prints them
// (c0.c0.c0, c0.c0.c1, c0.c1.c0, ... c1.c2.c1).  Used by the key generator for alpha_g1_beta_g2.
namespace zkv {
void reduced_pairing(const HG1Affine &P, const HG2Affine &Q, HFq out[12]) {
    const Fq12 f = final_exponentiation(multi_miller({P}, {Q}));
    const HFq2 *c[6] = {&f.c0.c0, &f.c0.c1, &f.c0.c2, &f.c1.c0, &f.c1.c1, &f.c1.c2};
    for (int i = 0; i < 6; i++) { out[2 * i] = c[i]->c0; out[2 * i + 1] = c[i]->c1; }
}
} // namespace zkv

bool verifyMintproof(char *data, char *cmtA_old_string, char *sn_old_string, char *cmtA_string, uint64_t value_s) {
    uint8_t a[32], b[32], c[32];
    zkw::parse_hex_blob(cmtA_old_string, a, 32); zkw::parse_hex_blob(sn_old_string, b, 32); zkw::parse_hex_blob(cmtA_string, c, 32);
    std::vector<bool> bits; push_blob(bits, a, 32); push_blob(bits, b, 32); push_blob(bits, c, 32); push_u64(bits, value_s);
    const bool ok = verify(ZKB200_MINT, data, bits);
    printf(ok ? "Verifying mint proof successfully!!!\n" : "Verifying mint proof unsuccessfully!!!\n");
    return ok;
}
bool verifyRedeemproof(char *data, char *cmtA_old_string, char *sn_old_string, char *cmtA_string, uint64_t value_s) {
    uint8_t a[32], b[32], c[32];
    zkw::parse_hex_blob(cmtA_old_string, a, 32); zkw::parse_hex_blob(sn_old_string, b, 32); zkw::parse_hex_blob(cmtA_string, c, 32);
    std::vector<bool> bits; push_blob(bits, a, 32); push_blob(bits, b, 32); push_blob(bits, c, 32); push_u64(bits, value_s);
    const bool ok = verify(ZKB200_REDEEM, data, bits);
    printf(ok ? "Verifying redeem proof successfully!!!\n" : "Verifying redeem proof unsuccessfully!!!\n");
    return ok;
}
bool verifySendproof(char *data, char *cmtA_old_string, char *sn_old_string, char *cmtS_string, char *cmtA_new_string) {
    uint8_t a[32], b[32], c[32], d[32];
    zkw::parse_hex_blob(cmtA_old_string, a, 32); zkw::parse_hex_blob(sn_old_string, b, 32); zkw::parse_hex_blob(cmtS_string, c, 32);
    zkw::parse_hex_blob(cmtA_new_string, d, 32);
    std::vector<bool> bits; push_blob(bits, a, 32); push_blob(bits, b, 32); push_blob(bits, c, 32); push_blob(bits, d, 32);
    const bool ok = verify(ZKB200_SEND, data, bits);
    printf(ok ? "Verifying send proof successfully!!!\n" : "Verifying send proof unsuccessfully!!!\n");
    return ok;
}
bool verifyDepositproof(char *data, char *RT, char *pk, char *cmtb_old, char *snold, char *cmtb, char *sns) {
    uint8_t rt[32], p[20], a[32], b[32], c[32], d[32];
    zkw::parse_hex_blob(RT, rt, 32); zkw::parse_hex_blob(pk, p, 20); zkw::parse_hex_blob(cmtb_old, a, 32); zkw::parse_hex_blob(snold, b, 32);
    zkw::parse_hex_blob(cmtb, c, 32); zkw::parse_hex_blob(sns, d, 32);
    std::vector<bool> bits; push_blob(bits, rt, 32); push_blob(bits, p, 20); push_blob(bits, a, 32); push_blob(bits, b, 32); push_blob(bits, c, 32);
    push_blob(bits, d, 32);
    const bool ok = verify(ZKB200_DEPOSIT, data, bits);
    printf(ok ? "Verifying deposit proof successfully!!!\n" : "Verifying deposit proof unsuccessfully!!!\n");
    return ok;
}
