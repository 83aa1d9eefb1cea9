// Host-side BN254 arithmetic for the glue around the GPU kernels (proof assembly, the final Horner step of each MSM,
// point normalisation, Z^-1 constants of the evaluation domains, key parsing).  4 x 64-bit Montgomery limbs with
// unsigned __int128 -- the same representation (x * 2^256 mod p, little-endian) as the device code and as libff's
// mont_repr (libff/algebra/fields/fp.tcc), so buffers move between host, device and key files without conversion.
// Not a hot path: everything here runs O(1)..O(windows) times per proof.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace zkh {
typedef unsigned __int128 u128;

struct FrTag {
    static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
    static constexpr uint64_t ONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
    static constexpr uint64_t INV = 0xc2e1f593efffffffull;
};
struct FqTag {
    static constexpr uint64_t MOD[4] = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ull, 0xb5e71911d44501fbull, 0x47ab1eff0a417ff6ull, 0x06d89f71cab8351full};
    static constexpr uint64_t ONE[4] = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
    static constexpr uint64_t INV = 0x87d20782e4866389ull;
};

template <class T> struct HFp {
    uint64_t v[4];
    static HFp zero() { HFp r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0; return r; }
    static HFp one() { HFp r; memcpy(r.v, T::ONE, 32); return r; }
    static HFp raw(const uint64_t x[4]) { HFp r; memcpy(r.v, x, 32); return r; }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const HFp &o) const { return memcmp(v, o.v, 32) == 0; }
    bool operator!=(const HFp &o) const { return !(*this == o); }
    static bool geq_mod(const uint64_t x[4]) {
        for (int i = 3; i >= 0; i--) { if (x[i] > T::MOD[i]) return true; if (x[i] < T::MOD[i]) return false; }
        return true;
    }
    static void sub_mod(uint64_t x[4]) {
        u128 b = 0;
        for (int i = 0; i < 4; i++) { u128 d = (u128)x[i] - T::MOD[i] - (uint64_t)b; x[i] = (uint64_t)d; b = (d >> 64) & 1; }
    }
    // r = x - MOD if x >= MOD (branch-free select)
    static void cond_sub(uint64_t x[4]) {
        uint64_t d[4]; u128 b = 0;
        for (int i = 0; i < 4; i++) { const u128 t = (u128)x[i] - T::MOD[i] - (uint64_t)b; d[i] = (uint64_t)t; b = (t >> 64) & 1; }
        const uint64_t keep = (uint64_t)0 - (uint64_t)b;               // all ones if x < MOD
        for (int i = 0; i < 4; i++) x[i] = (x[i] & keep) | (d[i] & ~keep);
    }
    HFp operator+(const HFp &o) const {
        HFp r; u128 c = 0;
        for (int i = 0; i < 4; i++) { c += (u128)v[i] + o.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }      // no carry out: both < 2^254
        if (geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    HFp operator-(const HFp &o) const {
        HFp r; u128 b = 0;
        for (int i = 0; i < 4; i++) { u128 d = (u128)v[i] - o.v[i] - (uint64_t)b; r.v[i] = (uint64_t)d; b = (d >> 64) & 1; }
        if (b) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.v[i] + T::MOD[i]; r.v[i] = (uint64_t)c; c >>= 64; } }
        return r;
    }
    HFp neg() const { return is_zero() ? *this : zero() - *this; }
    HFp dbl() const { return *this + *this; }
    // Montgomery product, CIOS without the extra carry word (valid because the top limb of both moduli is < 2^62), fully unrolled
    HFp operator*(const HFp &o) const {
        const uint64_t *a = v, *b = o.v;
        uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0;
#define ZKH_ROW(bi)                                                                                                    \
        {                                                                                                              \
            u128 x = (u128)a[0] * (bi) + t0; uint64_t A = (uint64_t)(x >> 64);                                         \
            const uint64_t m = (uint64_t)x * T::INV;                                                                   \
            u128 y = (u128)m * T::MOD[0] + (uint64_t)x; uint64_t C = (uint64_t)(y >> 64);                              \
            x = (u128)a[1] * (bi) + t1 + A; A = (uint64_t)(x >> 64); y = (u128)m * T::MOD[1] + (uint64_t)x + C; C = (uint64_t)(y >> 64); t0 = (uint64_t)y; \
            x = (u128)a[2] * (bi) + t2 + A; A = (uint64_t)(x >> 64); y = (u128)m * T::MOD[2] + (uint64_t)x + C; C = (uint64_t)(y >> 64); t1 = (uint64_t)y; \
            x = (u128)a[3] * (bi) + t3 + A; A = (uint64_t)(x >> 64); y = (u128)m * T::MOD[3] + (uint64_t)x + C; C = (uint64_t)(y >> 64); t2 = (uint64_t)y; \
            t3 = C + A;                                                                                                \
        }
        ZKH_ROW(b[0]) ZKH_ROW(b[1]) ZKH_ROW(b[2]) ZKH_ROW(b[3])
#undef ZKH_ROW
        HFp r; r.v[0] = t0; r.v[1] = t1; r.v[2] = t2; r.v[3] = t3;
        cond_sub(r.v);
        return r;
    }
    HFp sqr() const { return *this * *this; }
    HFp pow(const uint64_t e[4]) const {
        HFp r = one();
        for (int i = 255; i >= 0; i--) { r = r.sqr(); if ((e[i >> 6] >> (i & 63)) & 1) r = r * *this; }
        return r;
    }
    HFp pow64(uint64_t e) const { uint64_t ee[4] = {e, 0, 0, 0}; return pow(ee); }
    HFp inverse() const {
        uint64_t e[4]; memcpy(e, T::MOD, 32); e[0] -= 2;        // MOD[0] >= 2, no borrow
        return pow(e);
    }
    // canonical integer <-> Montgomery
    static HFp from_canonical(const uint64_t x[4]) { return raw(x) * raw(T::R2); }
    static HFp from_u64(uint64_t x) { uint64_t t[4] = {x, 0, 0, 0}; return from_canonical(t); }
    void to_canonical(uint64_t out[4]) const { uint64_t o[4] = {1, 0, 0, 0}; HFp r = *this * raw(o); memcpy(out, r.v, 32); }
    // decimal string (canonical) -> element; returns false on bad input
    static bool from_dec(const char *s, size_t len, HFp &out) {
        uint64_t x[4] = {0, 0, 0, 0};
        for (size_t k = 0; k < len; k++) {
            if (s[k] < '0' || s[k] > '9') return false;
            u128 c = (u128)(s[k] - '0');
            for (int i = 0; i < 4; i++) { c += (u128)x[i] * 10; x[i] = (uint64_t)c; c >>= 64; }
            if (c) return false;
        }
        out = from_canonical(x);
        return true;
    }
};
typedef HFp<FrTag> HFr;
typedef HFp<FqTag> HFq;

struct HFq2 {
    HFq c0, c1;
    static HFq2 zero() { return HFq2{HFq::zero(), HFq::zero()}; }
    static HFq2 one() { return HFq2{HFq::one(), HFq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const HFq2 &o) const { return c0 == o.c0 && c1 == o.c1; }
    bool operator!=(const HFq2 &o) const { return !(*this == o); }
    HFq2 operator+(const HFq2 &o) const { return HFq2{c0 + o.c0, c1 + o.c1}; }
    HFq2 operator-(const HFq2 &o) const { return HFq2{c0 - o.c0, c1 - o.c1}; }
    HFq2 neg() const { return HFq2{c0.neg(), c1.neg()}; }
    HFq2 dbl() const { return HFq2{c0.dbl(), c1.dbl()}; }
    HFq2 operator*(const HFq2 &o) const {
        HFq aa = c0 * o.c0, bb = c1 * o.c1;
        return HFq2{aa - bb, (c0 + c1) * (o.c0 + o.c1) - aa - bb};
    }
    HFq2 sqr() const { HFq ab = c0 * c1; return HFq2{(c0 + c1) * (c0 - c1), ab.dbl()}; }
    HFq2 inverse() const { HFq t = (c0.sqr() + c1.sqr()).inverse(); return HFq2{c0 * t, (c1 * t).neg()}; }
};

// XYZZ points on the host, same formulas and memory layout as the device (ec.cuh)
template <class F> struct HAffine {
    F x, y;
    bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static HAffine inf() { return HAffine{F::zero(), F::zero()}; }
};
template <class F> struct HPoint {
    F X, Y, ZZ, ZZZ;
    static HPoint inf() { return HPoint{F::zero(), F::zero(), F::zero(), F::zero()}; }
    bool is_inf() const { return ZZ.is_zero(); }
    static HPoint from_affine(const HAffine<F> &a) { return a.is_inf() ? inf() : HPoint{a.x, a.y, F::one(), F::one()}; }
    HPoint neg() const { return HPoint{X, Y.neg(), ZZ, ZZZ}; }
    HPoint dbl() const {
        if (is_inf()) return *this;
        F U = Y.dbl(), V = U.sqr(), W = U * V, S = X * V, XX = X.sqr(), M = XX.dbl() + XX;
        HPoint r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * Y;
        r.ZZ = V * ZZ;
        r.ZZZ = W * ZZZ;
        return r;
    }
    HPoint add(const HPoint &o) const {
        if (o.is_inf()) return *this;
        if (is_inf()) return o;
        F U1 = X * o.ZZ, U2 = o.X * ZZ, S1 = Y * o.ZZZ, S2 = o.Y * ZZZ;
        F P = U2 - U1, R = S2 - S1;
        if (P.is_zero()) return R.is_zero() ? dbl() : inf();
        F PP = P.sqr(), PPP = P * PP, Q = U1 * PP;
        HPoint r;
        r.X = R.sqr() - PPP - Q.dbl();
        r.Y = R * (Q - r.X) - S1 * PPP;
        r.ZZ = ZZ * o.ZZ * PP;
        r.ZZZ = ZZZ * o.ZZZ * PPP;
        return r;
    }
    // k = canonical integer (4 x 64 little-endian); fixed 4-bit windows: 252 doublings + <= 64 additions
    HPoint mul(const uint64_t k[4]) const {
        HPoint tab[16];
        tab[0] = inf(); tab[1] = *this;
        for (int i = 2; i < 16; i++) tab[i] = (i & 1) ? tab[i - 1].add(*this) : tab[i >> 1].dbl();
        HPoint r = inf();
        bool started = false;
        for (int i = 63; i >= 0; i--) {
            const unsigned d = (unsigned)((k[i >> 4] >> ((i & 15) * 4)) & 15);
            if (started) { r = r.dbl().dbl().dbl().dbl(); }
            if (d) { r = started ? r.add(tab[d]) : tab[d]; started = true; }
        }
        return r;
    }
    HAffine<F> to_affine() const {
        if (is_inf()) return HAffine<F>::inf();
        F inv = (ZZ * ZZZ).inverse();
        return HAffine<F>{X * (inv * ZZZ), Y * (inv * ZZ)};
    }
};
typedef HAffine<HFq> HG1Affine;
typedef HAffine<HFq2> HG2Affine;
typedef HPoint<HFq> HG1;
typedef HPoint<HFq2> HG2;

} // namespace zkh
