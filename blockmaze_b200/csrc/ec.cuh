// Short-Weierstrass (a = 0) group law for alt_bn128 G1 (over Fq) and G2 (over Fq2), templated on the coordinate field.
//
// Replaces alt_bn128_G1/G2::{operator+, add, mixed_add, dbl} (libff/algebra/curves/alt_bn128/alt_bn128_g1.cpp:139-358,
// alt_bn128_g2.cpp:149-365; Jacobian coordinates).  Accumulators here use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ,
// ZZ^3 = ZZZ^2; EFD "xyzz" formulas): an accumulator += affine point costs 8M+2S instead of the Jacobian mixed add's
// 7M+4S, and no inversion is needed anywhere on the device.  Any correct group law gives the same affine result, which
// is all the reference's proof encoding exposes (mintcgo.cpp:130-175).
// Infinity: ZZ == 0 for XYZZ; an affine point is "infinity" when x == 0 && y == 0 (not on either curve since b != 0).
#pragma once
#include "ff.cuh"

// Group operations are deliberately NOT inlined on the device: one G2 addition is ~45 base-field multiplications
// (~15k SASS instructions); a call costs nothing next to that and keeps kernels small enough for the instruction cache.
#if defined(__CUDACC__)
#define ZK_EC __host__ __device__ __noinline__
#else
#define ZK_EC inline
#endif

namespace zk {

template <class F> struct Affine {
    F x, y;
    ZK_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    ZK_HD static Affine inf() { Affine a; a.x = F::zero(); a.y = F::zero(); return a; }
    ZK_HD Affine neg() const { Affine a; a.x = x; a.y = y.neg(); return a; }
};

template <class F> struct XYZZ {
    F X, Y, ZZ, ZZZ;
    ZK_HD bool is_inf() const { return ZZ.is_zero(); }
    ZK_HD static XYZZ inf() { XYZZ p; p.X = F::zero(); p.Y = F::zero(); p.ZZ = F::zero(); p.ZZZ = F::zero(); return p; }
    ZK_HD static XYZZ from_affine(const Affine<F> &a) {
        if (a.is_inf()) return inf();
        XYZZ p; p.X = a.x; p.Y = a.y; p.ZZ = F::one(); p.ZZZ = F::one(); return p;
    }
    ZK_HD XYZZ neg() const { XYZZ p = *this; p.Y = Y.neg(); return p; }

    // dbl-2008-s-1
    ZK_EC XYZZ dbl() const { return dbl_inl(); }
    // The doubling INSIDE add(): the operand travels by value.  Calling the member dbl() on *this from within the out-of-line add() read
    // out-of-bounds local memory on sm_100a (compute-sanitizer, scripts/ubench/team.cu: the nested call received a stale frame address for
    // `this`); equal summands are rare in an MSM tail but legal, so the path must not depend on that.
#if defined(__CUDACC__)
    static __host__ __device__ __noinline__ XYZZ dbl_value(const XYZZ p) { return p.dbl_inl(); }
#else
    static inline XYZZ dbl_value(const XYZZ p) { return p.dbl_inl(); }
#endif
    ZK_HD XYZZ dbl_inl() const {
        if (is_inf()) return *this;
        F U = Y.dbl(), V = U.sqr(), W = U * V, S = X * V;
        F XX = X.sqr(), M = XX.dbl() + XX;
        XYZZ r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * Y;
        r.ZZ = V * ZZ;
        r.ZZZ = W * ZZZ;
        return r;
    }
    // mdbl-2008-s-1: 2 * affine
    ZK_EC static XYZZ dbl_affine(const Affine<F> &a) {
        if (a.is_inf()) return inf();
        F U = a.y.dbl(), V = U.sqr(), W = U * V, S = a.x * V;
        F XX = a.x.sqr(), M = XX.dbl() + XX;
        XYZZ r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * a.y;
        r.ZZ = V;
        r.ZZZ = W;
        return r;
    }
    // madd-2008-s: this += affine
    ZK_EC void add_affine(const Affine<F> &a) { add_affine_inl(a); }
    ZK_HD void add_affine_inl(const Affine<F> &a) {
        if (a.is_inf()) return;
        if (is_inf()) { X = a.x; Y = a.y; ZZ = F::one(); ZZZ = F::one(); return; }
        F U2 = a.x * ZZ, S2 = a.y * ZZZ;
        F Pp = U2 - X, R = S2 - Y;
        if (Pp.is_zero()) {
            if (R.is_zero()) *this = dbl_affine(a); else *this = inf();
            return;
        }
        F PP = Pp.sqr(), PPP = Pp * PP, Q = X * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - Y * PPP;
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }
    // the same mixed addition with every field multiplication inlined (G1 accumulate kernel variant, see msm.cuh)
    // (field operations passed in: mul, sqr, and mul2(a, b, c, d) = a*b + c*d with one reduction)
    template <class M, class S, class M2> ZK_HD void add_affine_with(const Affine<F> &a, M mul, S sqr, M2 mul2) {
        if (a.is_inf()) return;
        if (is_inf()) { X = a.x; Y = a.y; ZZ = F::one(); ZZZ = F::one(); return; }
        F U2 = mul(a.x, ZZ), S2 = mul(a.y, ZZZ);
        F Pp = U2 - X, R = S2 - Y;
        if (Pp.is_zero()) {
            if (R.is_zero()) *this = dbl_affine(a); else *this = inf();
            return;
        }
        F PP = sqr(Pp), PPP = mul(Pp, PP), Q = mul(X, PP);
        F X3 = sqr(R) - PPP - Q.dbl();
        Y = mul2(R, Q - X3, F::modulus_minus(Y), PPP);              // R (Q - X3) - Y PPP
        X = X3;
        ZZ = mul(ZZ, PP);
        ZZZ = mul(ZZZ, PPP);
    }
    // add-2008-s: this += other
    ZK_EC void add(const XYZZ &o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        F U1 = X * o.ZZ, U2 = o.X * ZZ, S1 = Y * o.ZZZ, S2 = o.Y * ZZZ;
        F Pp = U2 - U1, R = S2 - S1;
        if (Pp.is_zero()) {
            if (R.is_zero()) { const XYZZ t = *this; *this = dbl_value(t); } else *this = inf();
            return;
        }
        F PP = Pp.sqr(), PPP = Pp * PP, Q = U1 * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - S1 * PPP;
        X = X3;
        ZZ = ZZ * o.ZZ * PP;
        ZZZ = ZZZ * o.ZZZ * PPP;
    }
    // k * this for a small non-negative integer k (double-and-add, MSB first)
    ZK_EC XYZZ mul_small(uint32_t k) const {
        XYZZ r = inf();
        int top = 31;
        while (top > 0 && !((k >> top) & 1)) top--;
        for (int b = top; b >= 0; b--) {
            r = r.dbl();
            if ((k >> b) & 1) r.add(*this);
        }
        return r;
    }
    // scalar as 8 little-endian limbs (plain integer)
    ZK_EC XYZZ mul(const uint32_t k[8]) const {
        XYZZ r = inf();
        for (int b = 255; b >= 0; b--) {
            r = r.dbl();
            if ((k[b >> 5] >> (b & 31)) & 1) r.add(*this);
        }
        return r;
    }
    // host-side normalisation (one field inversion)
    ZK_EC Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        // x = X/ZZ, y = Y/ZZZ ; 1/ZZ = ZZ^2 ... use 1/ZZZ and ZZ: 1/ZZ = ZZZ^2 / ZZ^4 is no cheaper; invert both via one inversion
        F zz_zzz = ZZ * ZZZ;
        F inv = zz_zzz.inverse();                     // (Fp: safegcd on the device, see ff.cuh inverse())
        Affine<F> a;
        a.x = X * (inv * ZZZ);
        a.y = Y * (inv * ZZ);
        return a;
    }
};

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq> G1XYZZ;
typedef XYZZ<Fq2> G2XYZZ;

} // namespace zk
