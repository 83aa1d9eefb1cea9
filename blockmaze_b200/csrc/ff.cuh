// BN254 (alt_bn128) prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// Replaces the reference's 4 x 64-bit Fp_model (libff/algebra/fields/fp.tcc:23-190 mul_reduce, :407-507 add/sub,
// fp_aux.tcc x86-64 asm) with carry-chained 32-bit IMAD sequences (mad.lo.cc / madc.hi.cc).  The Montgomery
// representative x*2^256 mod p is bit-identical to libff's mont_repr, so proving-key bytes are used as stored.
//
// The multiplication is an interleaved (CIOS) product-scanning scheme that keeps the running sum split into an
// "even" and an "odd" column accumulator so that every row of partial products is one uninterrupted carry chain
// of mad.lo.cc/madc.hi.cc pairs (ptxas fuses each pair into one IMAD.WIDE.U32.X): 2*8*8 wide multiply-adds per
// modmul plus 8 single multiplies for the Montgomery quotients.
//
// The same source compiles for the host (carry flag emulated in C) so the algorithm is unit-tested on CPU and the
// host-side glue (proof assembly, pairing) shares one implementation.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_D __device__ __forceinline__
#else
#define ZK_HD inline
#define ZK_D inline
#endif

namespace zk {

// ---------------------------------------------------------------------------------------------------------------
// carry-flag primitives: PTX on device, C emulation on host
struct Carry {
#if !defined(__CUDA_ARCH__)
    uint32_t cf = 0;
#endif
    // d = a + b, set carry
    ZK_HD void add_cc(uint32_t &d, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
#else
        uint64_t s = (uint64_t)a + b; d = (uint32_t)s; cf = (uint32_t)(s >> 32);
#endif
    }
    ZK_HD void addc_cc(uint32_t &d, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
#else
        uint64_t s = (uint64_t)a + b + cf; d = (uint32_t)s; cf = (uint32_t)(s >> 32);
#endif
    }
    ZK_HD void addc(uint32_t &d, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
#else
        d = a + b + cf;
#endif
    }
    ZK_HD void sub_cc(uint32_t &d, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
#else
        uint64_t s = (uint64_t)a - b; d = (uint32_t)s; cf = (uint32_t)((s >> 32) & 1);   // cf = borrow
#endif
    }
    ZK_HD void subc_cc(uint32_t &d, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
#else
        uint64_t s = (uint64_t)a - b - cf; d = (uint32_t)s; cf = (uint32_t)((s >> 32) & 1);
#endif
    }
    // d = 0 - 0 - borrow  -> 0 or 0xffffffff
    ZK_HD void subc(uint32_t &d, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
#else
        d = a - b - cf;
#endif
    }
    // lo/hi(a*b) + c with carry handling
    ZK_HD void mad_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
        asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
#else
        uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c; d = (uint32_t)s; cf = (uint32_t)(s >> 32);
#endif
    }
    ZK_HD void madc_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
#else
        uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c + cf; d = (uint32_t)s; cf = (uint32_t)(s >> 32);
#endif
    }
    ZK_HD void madc_hi_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
#else
        uint64_t s = (((uint64_t)a * b) >> 32) + c + cf; d = (uint32_t)s; cf = (uint32_t)(s >> 32);
#endif
    }
    ZK_HD void madc_hi(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
        asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
#else
        d = (uint32_t)(((uint64_t)a * b) >> 32) + c + cf;
#endif
    }
};

ZK_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
ZK_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// field parameter packs (alt_bn128_init.cpp:96-144).  Limbs little-endian.
struct FrParams {   // scalar field r
    static constexpr uint32_t INV = 0xefffffffu;   // -r^-1 mod 2^32
    ZK_HD static constexpr uint32_t mod(int i) {
        return i == 0 ? 0xf0000001u : i == 1 ? 0x43e1f593u : i == 2 ? 0x79b97091u : i == 3 ? 0x2833e848u :
               i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
    ZK_HD static constexpr uint32_t one(int i) {     // R mod r
        return i == 0 ? 0x4ffffffbu : i == 1 ? 0xac96341cu : i == 2 ? 0x9f60cd29u : i == 3 ? 0x36fc7695u :
               i == 4 ? 0x7879462eu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
    }
    ZK_HD static constexpr uint32_t r2(int i) {      // R^2 mod r
        return i == 0 ? 0xae216da7u : i == 1 ? 0x1bb8e645u : i == 2 ? 0xe35c59e3u : i == 3 ? 0x53fe3ab1u :
               i == 4 ? 0x53bb8085u : i == 5 ? 0x8c49833du : i == 6 ? 0x7f4e44a5u : 0x0216d0b1u;
    }
    ZK_HD static constexpr uint32_t r3(int i) {      // R^3 mod r (turns the plain inverse of a Montgomery representative back into Montgomery form)
        return i == 0 ? 0xb4bf0040u : i == 1 ? 0x5e94d8e1u : i == 2 ? 0x1cfbb6b8u : i == 3 ? 0x2a489cbeu :
               i == 4 ? 0xa19fcfedu : i == 5 ? 0x893cc664u : i == 6 ? 0x7fcc657cu : 0x0cf8594bu;
    }
};
struct FqParams {   // base field q
    static constexpr uint32_t INV = 0xe4866389u;   // -q^-1 mod 2^32
    ZK_HD static constexpr uint32_t mod(int i) {
        return i == 0 ? 0xd87cfd47u : i == 1 ? 0x3c208c16u : i == 2 ? 0x6871ca8du : i == 3 ? 0x97816a91u :
               i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
    ZK_HD static constexpr uint32_t one(int i) {     // R mod q
        return i == 0 ? 0xc58f0d9du : i == 1 ? 0xd35d438du : i == 2 ? 0xf5c70b3du : i == 3 ? 0x0a78eb28u :
               i == 4 ? 0x7879462cu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
    }
    ZK_HD static constexpr uint32_t r2(int i) {      // R^2 mod q
        return i == 0 ? 0x538afa89u : i == 1 ? 0xf32cfc5bu : i == 2 ? 0xd44501fbu : i == 3 ? 0xb5e71911u :
               i == 4 ? 0x0a417ff6u : i == 5 ? 0x47ab1effu : i == 6 ? 0xcab8351fu : 0x06d89f71u;
    }
    ZK_HD static constexpr uint32_t r3(int i) {      // R^3 mod q
        return i == 0 ? 0xda1530dfu : i == 1 ? 0xb1cd6dafu : i == 2 ? 0xa7283db6u : i == 3 ? 0x62f210e6u :
               i == 4 ? 0x0ada0afbu : i == 5 ? 0xef7f0b0cu : i == 6 ? 0x2d592544u : 0x20fd6e90u;
    }
};

// ---------------------------------------------------------------------------------------------------------------
template <class P> struct Fp {
    uint32_t v[8];

    ZK_HD static Fp zero() { Fp r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
    ZK_HD static Fp one() { Fp r; for (int i = 0; i < 8; i++) r.v[i] = P::one(i); return r; }
    ZK_HD static Fp r2() { Fp r; for (int i = 0; i < 8; i++) r.v[i] = P::r2(i); return r; }
    ZK_HD static Fp modulus() { Fp r; for (int i = 0; i < 8; i++) r.v[i] = P::mod(i); return r; }

    ZK_HD bool is_zero() const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= v[i]; return o == 0; }
    ZK_HD bool operator==(const Fp &b) const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= v[i] ^ b.v[i]; return o == 0; }
    ZK_HD bool operator!=(const Fp &b) const { return !(*this == b); }

    // r = a - p if a >= p (input < 2p)
    ZK_HD void reduce_once() {
        Carry c; uint32_t t[8], brw;
        c.sub_cc(t[0], v[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < 8; i++) c.subc_cc(t[i], v[i], P::mod(i));
        c.subc(brw, 0, 0);
        if (brw == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = t[i];
        }
    }

    ZK_HD friend Fp operator+(const Fp &a, const Fp &b) {
        Fp r; Carry c;
        c.add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) c.addc_cc(r.v[i], a.v[i], b.v[i]);   // no overflow: a,b < p < 2^254
        r.reduce_once();
        return r;
    }
    ZK_HD friend Fp operator-(const Fp &a, const Fp &b) {
        Fp r; Carry c; uint32_t brw;
        c.sub_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) c.subc_cc(r.v[i], a.v[i], b.v[i]);
        c.subc(brw, 0, 0);            // 0xffffffff if a < b
        Carry d;
        d.add_cc(r.v[0], r.v[0], P::mod(0) & brw);
#pragma unroll
        for (int i = 1; i < 8; i++) d.addc_cc(r.v[i], r.v[i], P::mod(i) & brw);
        return r;
    }
    ZK_HD Fp neg() const { return is_zero() ? *this : modulus_minus(*this); }
    ZK_HD static Fp modulus_minus(const Fp &a) {
        Fp r; Carry c;
        c.sub_cc(r.v[0], P::mod(0), a.v[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) c.subc_cc(r.v[i], P::mod(i), a.v[i]);
        return r;
    }
    ZK_HD Fp dbl() const { return *this + *this; }

    // Montgomery product a*b*R^-1 mod p, inputs and output fully reduced (< p).
    //
    // acc[0] ("E") holds 64-bit partial products whose low limb sits at an EVEN absolute column, acc[1] ("O") those at an
    // ODD column; the running sum is E + O.  Row i adds a*b_i and m_i*p at columns i..i+8, so nothing is ever shifted:
    // every multiply-add has destination == addend and ptxas fuses each mad.lo.cc/madc.hi.cc pair into one
    // IMAD.WIDE.U32.X.  H = acc[i&1] owns column i ("hot"), C = the other one has a dangling high limb there, which is
    // merged first (add.cc) so that the quotient digit m_i sees the whole column; its carry rides into C's chain.
    ZK_HD friend Fp operator*(const Fp &a, const Fp &b) {
#if defined(__CUDA_ARCH__) && !defined(ZK_INLINE_MUL)
        return mul_call(a, b);
#else
        return mul_impl(a, b);
#endif
    }
#if defined(__CUDACC__)
    // One shared copy of the ~360-instruction multiplication per field: group-law kernels call it 10-45 times per point
    // operation, and keeping it out of line keeps their hot loop inside the instruction cache (ncu: the inlined version of
    // msm_accumulate_kernel stalled mostly on no_instruction, icc hit rate 74 %).
    static __device__ __noinline__ Fp mul_call(const Fp a, const Fp b) { return mul_impl(a, b); }
#endif
    ZK_HD static Fp mul_impl(const Fp &a, const Fp &b) { return mul_core<true>(a, b); }
    // ---- lazy arithmetic for the NTT butterflies (ntt.cuh): values live in [0, 4p), 4p < 2^256 for both BN254 moduli -----------------------
    // a < 4p (any 256-bit value whose top limb stays below 2^32 - 8), b < p  ->  a*b*R^-1 + (0 or p), below 2p, NOT reduced
    ZK_HD static Fp mul_lazy(const Fp &a, const Fp &b) { return mul_core<false>(a, b); }
    ZK_HD static constexpr uint32_t mod2(int i) { return i == 0 ? P::mod(0) << 1 : (P::mod(i) << 1) | (P::mod(i - 1) >> 31); }      // limb i of 2p
    // a < 4p  ->  a or a - 2p, below 2p
    ZK_HD static Fp condsub_2p(const Fp &a) {
        Carry c; Fp t; uint32_t brw;
        c.sub_cc(t.v[0], a.v[0], mod2(0));
#pragma unroll
        for (int i = 1; i < 8; i++) c.subc_cc(t.v[i], a.v[i], mod2(i));
        c.subc(brw, 0, 0);
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = brw ? a.v[i] : t.v[i];
        return r;
    }
    // a, b < 2p  ->  a + b, below 4p
    ZK_HD static Fp add_lazy(const Fp &a, const Fp &b) {
        Fp r; Carry c;
        c.add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) c.addc_cc(r.v[i], a.v[i], b.v[i]);
        c.addc(r.v[7], a.v[7], b.v[7]);
        return r;
    }
    // a, b < 2p  ->  a - b + 2p, in (0, 4p)
    ZK_HD static Fp sub_lazy(const Fp &a, const Fp &b) {
        Fp r; Carry c;
        c.add_cc(r.v[0], a.v[0], mod2(0));
#pragma unroll
        for (int i = 1; i < 7; i++) c.addc_cc(r.v[i], a.v[i], mod2(i));
        c.addc(r.v[7], a.v[7], mod2(7));
        Carry d;
        d.sub_cc(r.v[0], r.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) d.subc_cc(r.v[i], r.v[i], b.v[i]);
        d.subc(r.v[7], r.v[7], b.v[7]);
        return r;
    }
    // a < 4p  ->  a mod p, fully reduced
    ZK_HD static Fp reduce_4p(const Fp &a) { Fp r = condsub_2p(a); r.reduce_once(); return r; }
    template <bool REDUCE> ZK_HD static Fp mul_core(const Fp &a, const Fp &b) {
        uint32_t acc[2][18];
#pragma unroll
        for (int k = 0; k < 18; k++) { acc[0][k] = 0; acc[1][k] = 0; }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t *H = acc[i & 1], *Cc = acc[(i & 1) ^ 1];
            const uint32_t bi = b.v[i];
            Carry c;
            if (i == 0) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    H[j] = mul_lo(a.v[j], bi);      H[j + 1] = mul_hi(a.v[j], bi);
                    Cc[j + 1] = mul_lo(a.v[j + 1], bi); Cc[j + 2] = mul_hi(a.v[j + 1], bi);
                }
            } else {
                c.add_cc(H[i], H[i], Cc[i]);
#pragma unroll
                for (int j = 1; j < 8; j += 2) {
                    c.madc_lo_cc(Cc[i + j], a.v[j], bi, Cc[i + j]);
                    if (j < 7) c.madc_hi_cc(Cc[i + j + 1], a.v[j], bi, Cc[i + j + 1]);
                    else c.madc_hi(Cc[i + j + 1], a.v[j], bi, Cc[i + j + 1]);
                }
                Carry d;
                d.mad_lo_cc(H[i], a.v[0], bi, H[i]);
                d.madc_hi_cc(H[i + 1], a.v[0], bi, H[i + 1]);
#pragma unroll
                for (int j = 2; j < 8; j += 2) {
                    d.madc_lo_cc(H[i + j], a.v[j], bi, H[i + j]);
                    d.madc_hi_cc(H[i + j + 1], a.v[j], bi, H[i + j + 1]);
                }
                d.addc(Cc[i + 8], Cc[i + 8], 0);
            }
            const uint32_t m = mul_lo(H[i], P::INV);
            Carry e;
            e.mad_lo_cc(Cc[i + 1], P::mod(1), m, Cc[i + 1]);
            e.madc_hi_cc(Cc[i + 2], P::mod(1), m, Cc[i + 2]);
#pragma unroll
            for (int j = 3; j < 8; j += 2) {
                e.madc_lo_cc(Cc[i + j], P::mod(j), m, Cc[i + j]);
                if (j < 7) e.madc_hi_cc(Cc[i + j + 1], P::mod(j), m, Cc[i + j + 1]);
                else e.madc_hi(Cc[i + j + 1], P::mod(j), m, Cc[i + j + 1]);
            }
            Carry f;
            f.mad_lo_cc(H[i], P::mod(0), m, H[i]);
            f.madc_hi_cc(H[i + 1], P::mod(0), m, H[i + 1]);
#pragma unroll
            for (int j = 2; j < 8; j += 2) {
                f.madc_lo_cc(H[i + j], P::mod(j), m, H[i + j]);
                f.madc_hi_cc(H[i + j + 1], P::mod(j), m, H[i + j + 1]);
            }
            f.addc(Cc[i + 8], Cc[i + 8], 0);
        }
        // after row 7 (H = O, C = E): result = columns 8..15 of E + O
        Fp r; Carry c;
        c.add_cc(r.v[0], acc[0][8], acc[1][8]);
#pragma unroll
        for (int k = 1; k < 7; k++) c.addc_cc(r.v[k], acc[0][8 + k], acc[1][8 + k]);
        c.addc(r.v[7], acc[0][15], 0);
        if (REDUCE) r.reduce_once();
        return r;
    }
    // (a*b + c*d) * R^-1 mod p with ONE Montgomery reduction: every row adds a*b_i, c*d_i and m_i*p to the same two accumulators, 3*64 + 8
    // multiplications instead of 2*(2*64 + 8).  a*b + c*d < 2 p^2 keeps the quotient rows exact (2p < R) and the result below 2p; the column
    // where a row deposits its carries receives three high limbs below 2^30 each (a_7, c_7, p_7 < 2^30), so it still cannot overflow.
    // The XYZZ addition uses it for Y3 = R*(Q - X3) + (p - Y1)*PPP.
    ZK_HD static Fp mul2_impl(const Fp &a, const Fp &b, const Fp &c, const Fp &d) {
        uint32_t acc[2][18];
#pragma unroll
        for (int k = 0; k < 18; k++) { acc[0][k] = 0; acc[1][k] = 0; }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t *H = acc[i & 1], *Cc = acc[(i & 1) ^ 1];
            const uint32_t bi = b.v[i], di = d.v[i];
            Carry k1;
            if (i > 0) k1.add_cc(H[i], H[i], Cc[i]);
            // a*b_i: odd j into Cc (chain k1 carries the merge above), even j into H
#pragma unroll
            for (int j = 1; j < 8; j += 2) {
                if (i == 0 && j == 1) k1.mad_lo_cc(Cc[i + j], a.v[j], bi, Cc[i + j]); else k1.madc_lo_cc(Cc[i + j], a.v[j], bi, Cc[i + j]);
                if (j < 7) k1.madc_hi_cc(Cc[i + j + 1], a.v[j], bi, Cc[i + j + 1]);
                else k1.madc_hi(Cc[i + j + 1], a.v[j], bi, Cc[i + j + 1]);
            }
            Carry k2;
            k2.mad_lo_cc(H[i], a.v[0], bi, H[i]);
            k2.madc_hi_cc(H[i + 1], a.v[0], bi, H[i + 1]);
#pragma unroll
            for (int j = 2; j < 8; j += 2) {
                k2.madc_lo_cc(H[i + j], a.v[j], bi, H[i + j]);
                k2.madc_hi_cc(H[i + j + 1], a.v[j], bi, H[i + j + 1]);
            }
            k2.addc(Cc[i + 8], Cc[i + 8], 0);
            // c*d_i
            Carry k3;
            k3.mad_lo_cc(Cc[i + 1], c.v[1], di, Cc[i + 1]);
            k3.madc_hi_cc(Cc[i + 2], c.v[1], di, Cc[i + 2]);
#pragma unroll
            for (int j = 3; j < 8; j += 2) {
                k3.madc_lo_cc(Cc[i + j], c.v[j], di, Cc[i + j]);
                if (j < 7) k3.madc_hi_cc(Cc[i + j + 1], c.v[j], di, Cc[i + j + 1]);
                else k3.madc_hi(Cc[i + j + 1], c.v[j], di, Cc[i + j + 1]);
            }
            Carry k4;
            k4.mad_lo_cc(H[i], c.v[0], di, H[i]);
            k4.madc_hi_cc(H[i + 1], c.v[0], di, H[i + 1]);
#pragma unroll
            for (int j = 2; j < 8; j += 2) {
                k4.madc_lo_cc(H[i + j], c.v[j], di, H[i + j]);
                k4.madc_hi_cc(H[i + j + 1], c.v[j], di, H[i + j + 1]);
            }
            k4.addc(Cc[i + 8], Cc[i + 8], 0);
            // m_i * p
            const uint32_t m = mul_lo(H[i], P::INV);
            Carry e;
            e.mad_lo_cc(Cc[i + 1], P::mod(1), m, Cc[i + 1]);
            e.madc_hi_cc(Cc[i + 2], P::mod(1), m, Cc[i + 2]);
#pragma unroll
            for (int j = 3; j < 8; j += 2) {
                e.madc_lo_cc(Cc[i + j], P::mod(j), m, Cc[i + j]);
                if (j < 7) e.madc_hi_cc(Cc[i + j + 1], P::mod(j), m, Cc[i + j + 1]);
                else e.madc_hi(Cc[i + j + 1], P::mod(j), m, Cc[i + j + 1]);
            }
            Carry f;
            f.mad_lo_cc(H[i], P::mod(0), m, H[i]);
            f.madc_hi_cc(H[i + 1], P::mod(0), m, H[i + 1]);
#pragma unroll
            for (int j = 2; j < 8; j += 2) {
                f.madc_lo_cc(H[i + j], P::mod(j), m, H[i + j]);
                f.madc_hi_cc(H[i + j + 1], P::mod(j), m, H[i + j + 1]);
            }
            f.addc(Cc[i + 8], Cc[i + 8], 0);
        }
        Fp r; Carry cz;
        cz.add_cc(r.v[0], acc[0][8], acc[1][8]);
#pragma unroll
        for (int k = 1; k < 7; k++) cz.addc_cc(r.v[k], acc[0][8 + k], acc[1][8 + k]);
        cz.addc(r.v[7], acc[0][15], acc[1][15]);
        r.reduce_once();
        return r;
    }
    // Montgomery square a*a*R^-1 mod p with 36 + 64 wide multiply-adds instead of 64 + 64 (a_i*a_j = a_j*a_i is computed once).
    //   1. off-diagonal products a_i*a_j (i < j) go row by row into an even- and an odd-column accumulator like the rows of mul_impl; a chain
    //      that ends deposits its carry in the next column, which no earlier row has touched (row i reaches column i+8 at most);
    //   2. t = 2*(E + O) + sum a_i^2 * 2^(64 i): the diagonal products sit on disjoint column pairs, so they are ONE 16-limb carry chain;
    //   3. Montgomery reduction of the 16 limbs: the m_i*p half-rows of mul_impl on accumulators that start from t[0..8) and take in t[i+8]
    //      at the END of row i -- the columns where a row deposits a carry (i+8 of either accumulator) are then still small, as in mul_impl.
    ZK_HD Fp sqr() const {
#if defined(__CUDA_ARCH__) && !defined(ZK_INLINE_MUL)
        return sqr_call(*this);
#else
        return sqr_impl(*this);
#endif
    }
#if defined(__CUDACC__)
    static __device__ __noinline__ Fp sqr_call(const Fp a) { return sqr_impl(a); }
#endif
    ZK_HD static Fp sqr_impl(const Fp &a) {
        uint32_t off[2][16];                       // off[0]: products whose low limb sits at an even column, off[1]: odd
#pragma unroll
        for (int k = 0; k < 16; k++) { off[0][k] = 0; off[1][k] = 0; }
#pragma unroll
        for (int i = 0; i < 7; i++) {
#pragma unroll
            for (int par = 1; par <= 2; par++) {   // chain of the products a_i*a_j with j = i+par, i+par+2, ...
                if (i + par > 7) continue;
                uint32_t *X = off[(2 * i + par) & 1];
                Carry c;
                int last = 0;
#pragma unroll
                for (int j = i + par; j < 8; j += 2) {
                    if (j == i + par) c.mad_lo_cc(X[i + j], a.v[i], a.v[j], X[i + j]); else c.madc_lo_cc(X[i + j], a.v[i], a.v[j], X[i + j]);
                    c.madc_hi_cc(X[i + j + 1], a.v[i], a.v[j], X[i + j + 1]);
                    last = i + j + 1;
                }
                if (last + 1 < 16) c.addc(X[last + 1], X[last + 1], 0);
            }
        }
        uint32_t t[16];
        {   // t = (E + O), then doubled
            Carry c;
            c.add_cc(t[0], off[0][0], off[1][0]);
#pragma unroll
            for (int k = 1; k < 15; k++) c.addc_cc(t[k], off[0][k], off[1][k]);
            c.addc(t[15], off[0][15], off[1][15]);
            Carry d;
            d.add_cc(t[0], t[0], t[0]);
#pragma unroll
            for (int k = 1; k < 15; k++) d.addc_cc(t[k], t[k], t[k]);
            d.addc(t[15], t[15], t[15]);
            Carry e;                               // + the diagonal
            e.mad_lo_cc(t[0], a.v[0], a.v[0], t[0]);
            e.madc_hi_cc(t[1], a.v[0], a.v[0], t[1]);
#pragma unroll
            for (int i = 1; i < 8; i++) {
                e.madc_lo_cc(t[2 * i], a.v[i], a.v[i], t[2 * i]);
                if (i < 7) e.madc_hi_cc(t[2 * i + 1], a.v[i], a.v[i], t[2 * i + 1]); else e.madc_hi(t[2 * i + 1], a.v[i], a.v[i], t[2 * i + 1]);
            }
        }
        uint32_t acc[2][18];
#pragma unroll
        for (int k = 0; k < 18; k++) { acc[0][k] = k < 8 ? t[k] : 0; acc[1][k] = 0; }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t *H = acc[i & 1], *Cc = acc[(i & 1) ^ 1];
            Carry e;
            if (i > 0) e.add_cc(H[i], H[i], Cc[i]);            // the whole column i, its carry rides into the chain below
            const uint32_t m = mul_lo(H[i], P::INV);
            if (i > 0) e.madc_lo_cc(Cc[i + 1], P::mod(1), m, Cc[i + 1]); else e.mad_lo_cc(Cc[i + 1], P::mod(1), m, Cc[i + 1]);
            e.madc_hi_cc(Cc[i + 2], P::mod(1), m, Cc[i + 2]);
#pragma unroll
            for (int j = 3; j < 8; j += 2) {
                e.madc_lo_cc(Cc[i + j], P::mod(j), m, Cc[i + j]);
                if (j < 7) e.madc_hi_cc(Cc[i + j + 1], P::mod(j), m, Cc[i + j + 1]);
                else e.madc_hi(Cc[i + j + 1], P::mod(j), m, Cc[i + j + 1]);
            }
            Carry f;
            f.mad_lo_cc(H[i], P::mod(0), m, H[i]);
            f.madc_hi_cc(H[i + 1], P::mod(0), m, H[i + 1]);
#pragma unroll
            for (int j = 2; j < 8; j += 2) {
                f.madc_lo_cc(H[i + j], P::mod(j), m, H[i + j]);
                f.madc_hi_cc(H[i + j + 1], P::mod(j), m, H[i + j + 1]);
            }
            f.addc(Cc[i + 8], Cc[i + 8], 0);
            Carry g;                                           // now take in limb i+8 of the square (column i+9 is untouched so far)
            g.add_cc(acc[0][i + 8], acc[0][i + 8], t[i + 8]);
            g.addc(acc[0][i + 9], acc[0][i + 9], 0);
        }
        Fp r; Carry c;
        c.add_cc(r.v[0], acc[0][8], acc[1][8]);
#pragma unroll
        for (int k = 1; k < 7; k++) c.addc_cc(r.v[k], acc[0][8 + k], acc[1][8 + k]);
        c.addc(r.v[7], acc[0][15], acc[1][15]);
        r.reduce_once();
        return r;
    }

    // canonical <-> Montgomery
    ZK_HD Fp to_mont() const { return *this * r2(); }
    ZK_HD Fp from_mont() const { Fp o = zero(); o.v[0] = 1; return *this * o; }

    // a^e for a little-endian 8-limb exponent (plain integer), square-and-multiply MSB first
    ZK_HD Fp pow(const uint32_t e[8]) const {
        Fp r = one(); bool started = false;
        for (int i = 255; i >= 0; i--) {
            if (started) r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) { r = started ? r * *this : *this; started = true; }
        }
        return r;
    }
    ZK_HD Fp pow_u64(uint64_t e) const {
        uint32_t ee[8] = {(uint32_t)e, (uint32_t)(e >> 32), 0, 0, 0, 0, 0, 0};
        return pow(ee);
    }
    // Inverse by the Bernstein-Yang "safegcd" division steps (the half-delta variant with 30-bit signed limbs, as in libsecp256k1's
    // modinv32): 20 batches of 30 branch-free division steps on the low words of (f, g) = (p, x), each batch followed by one update of
    // the full-width (f, g) and of the Bezout pair (d, e) by the batch's 2x2 transition matrix.  About 16 k instructions, nearly all of
    // them 32-bit ALU work: 1/25 of the multiply-pipe time of the Fermat inverse below (254 squarings + ~127 multiplications), which is
    // what makes a per-thread batched inversion affordable inside the multiply-bound MSM kernels (msm.cuh, affine rounds).
    // Input and output are Montgomery representatives; 0 -> 0.  The reference inverts with GMP's mpn_gcdext (fp.tcc:650-679); any correct
    // inverse is the same field element.
    ZK_HD Fp inverse_gcd() const {
        const int32_t M30 = (int32_t)0x3fffffff;
        // p and x in nine signed 30-bit limbs
        int32_t f[9], g[9], m[9], d[9], e[9];
#pragma unroll
        for (int i = 0; i < 9; i++) {
            const int bit = 30 * i, w = bit >> 5, off = bit & 31;
            uint32_t pm = P::mod(w < 8 ? w : 7) >> off, xv = v[w < 8 ? w : 7] >> off;
            if (w >= 8) { pm = 0; xv = 0; }
            if (off > 2 && w + 1 < 8) { pm |= P::mod(w + 1) << (32 - off); xv |= v[w + 1] << (32 - off); }
            m[i] = (int32_t)(pm & (uint32_t)M30); f[i] = m[i]; g[i] = (int32_t)(xv & (uint32_t)M30);
            d[i] = 0; e[i] = 0;
        }
        e[0] = 1;
        const uint32_t minv30 = (0u - P::INV) & (uint32_t)M30;          // p^-1 mod 2^30
        int32_t zeta = -1;                                               // -(delta + 1/2), delta = 1/2
        for (int batch = 0; batch < 20; batch++) {
            // 30 division steps on the low limbs; the matrix (u v; q r) accumulates with a factor 2^30
            uint32_t u = 1, vv = 0, q = 0, r = 1;
            uint32_t ff = (uint32_t)f[0], gg = (uint32_t)g[0];
            for (int i = 0; i < 30; i++) {
                uint32_t c1 = (uint32_t)(zeta >> 31), c2 = 0u - (gg & 1u);
                const uint32_t x = (ff ^ c1) - c1, y = (u ^ c1) - c1, z = (vv ^ c1) - c1;
                gg += x & c2; q += y & c2; r += z & c2;
                c1 &= c2;
                zeta = (int32_t)(((uint32_t)zeta ^ c1) - 1u);
                ff += gg & c1; u += q & c1; vv += r & c1;
                gg >>= 1; u <<= 1; vv <<= 1;
            }
            const int32_t tu = (int32_t)u, tv = (int32_t)vv, tq = (int32_t)q, tr = (int32_t)r;
            {   // (d, e) <- (u d + v e, q d + r e) / 2^30 mod p, both kept in (-2p, p)
                const int32_t sd = d[8] >> 31, se = e[8] >> 31;
                int32_t md = (tu & sd) + (tv & se), me = (tq & sd) + (tr & se);
                int64_t cd = (int64_t)tu * d[0] + (int64_t)tv * e[0], ce = (int64_t)tq * d[0] + (int64_t)tr * e[0];
                md -= (int32_t)((minv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
                me -= (int32_t)((minv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
                cd += (int64_t)m[0] * md; ce += (int64_t)m[0] * me;
                cd >>= 30; ce >>= 30;
#pragma unroll
                for (int i = 1; i < 9; i++) {
                    cd += (int64_t)tu * d[i] + (int64_t)tv * e[i]; ce += (int64_t)tq * d[i] + (int64_t)tr * e[i];
                    cd += (int64_t)m[i] * md; ce += (int64_t)m[i] * me;
                    d[i - 1] = (int32_t)cd & M30; cd >>= 30; e[i - 1] = (int32_t)ce & M30; ce >>= 30;
                }
                d[8] = (int32_t)cd; e[8] = (int32_t)ce;
            }
            {   // (f, g) <- (u f + v g, q f + r g) / 2^30, exact
                int64_t cf = (int64_t)tu * f[0] + (int64_t)tv * g[0], cg = (int64_t)tq * f[0] + (int64_t)tr * g[0];
                cf >>= 30; cg >>= 30;
#pragma unroll
                for (int i = 1; i < 9; i++) {
                    cf += (int64_t)tu * f[i] + (int64_t)tv * g[i]; cg += (int64_t)tq * f[i] + (int64_t)tr * g[i];
                    f[i - 1] = (int32_t)cf & M30; cf >>= 30; g[i - 1] = (int32_t)cg & M30; cg >>= 30;
                }
                f[8] = (int32_t)cf; g[8] = (int32_t)cg;
            }
        }
        // now g = 0 and f = +-gcd = +-1 (or +-p for x = 0, in which case d = 0): x^-1 = sign(f) * d, brought into [0, p)
        {
            int32_t add = d[8] >> 31;
#pragma unroll
            for (int i = 0; i < 9; i++) d[i] += m[i] & add;
            const int32_t neg = f[8] >> 31;
#pragma unroll
            for (int i = 0; i < 9; i++) d[i] = (d[i] ^ neg) - neg;
#pragma unroll
            for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
            add = d[8] >> 31;
#pragma unroll
            for (int i = 0; i < 9; i++) d[i] += m[i] & add;
#pragma unroll
            for (int i = 0; i < 8; i++) { d[i + 1] += d[i] >> 30; d[i] &= M30; }
        }
        Fp o;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const int bit = 32 * w, i = bit / 30, off = bit % 30;
            uint32_t x = (uint32_t)d[i] >> off;
            if (i + 1 < 9) x |= (uint32_t)d[i + 1] << (30 - off);
            if (off > 28 && i + 2 < 9) x |= (uint32_t)d[i + 2] << (60 - off);
            o.v[w] = x;
        }
        Fp r3c; for (int i = 0; i < 8; i++) r3c.v[i] = P::r3(i);
        return o * r3c;                                                  // (xR)^-1 * R^3 / R = x^-1 R
    }
    // 1/a, 0 for 0: safegcd on the device (key load does 15-31 normalisations per base, keygen one per point); the host keeps the Fermat
    // chain, which its tests pin against Python
    ZK_HD Fp inverse() const {
#if defined(__CUDA_ARCH__)
        return inverse_gcd();
#else
        return inverse_fermat();
#endif
    }
    // Fermat inverse a^(p-2); returns 0 for 0
    ZK_HD Fp inverse_fermat() const {
        uint32_t e[8]; Carry c;
        c.sub_cc(e[0], P::mod(0), 2);
        for (int i = 1; i < 8; i++) c.subc_cc(e[i], P::mod(i), 0);
        return pow(e);
    }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// ---------------------------------------------------------------------------------------------------------------
// Fq2 = Fq[u]/(u^2 + 1)   (libff/algebra/fields/fp2.tcc; non_residue = -1, alt_bn128_init.cpp:151)
struct Fq2 {
    Fq c0, c1;
    ZK_HD static Fq2 zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
    ZK_HD static Fq2 one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
    ZK_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    ZK_HD bool operator==(const Fq2 &b) const { return c0 == b.c0 && c1 == b.c1; }
    ZK_HD bool operator!=(const Fq2 &b) const { return !(*this == b); }
    ZK_HD friend Fq2 operator+(const Fq2 &a, const Fq2 &b) { Fq2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
    ZK_HD friend Fq2 operator-(const Fq2 &a, const Fq2 &b) { Fq2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
    ZK_HD Fq2 neg() const { Fq2 r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
    ZK_HD Fq2 dbl() const { Fq2 r; r.c0 = c0.dbl(); r.c1 = c1.dbl(); return r; }
    ZK_HD friend Fq2 operator*(const Fq2 &a, const Fq2 &b) {      // Karatsuba, 3 base multiplications
        Fq aa = a.c0 * b.c0, bb = a.c1 * b.c1;
        Fq2 r;
        r.c1 = (a.c0 + a.c1) * (b.c0 + b.c1) - aa - bb;
        r.c0 = aa - bb;
        return r;
    }
    ZK_HD Fq2 sqr() const {                                         // complex squaring, 2 base multiplications
        Fq ab = c0 * c1;
        Fq2 r;
        r.c0 = (c0 + c1) * (c0 - c1);
        r.c1 = ab.dbl();
        return r;
    }
    ZK_HD Fq2 mul_fq(const Fq &s) const { Fq2 r; r.c0 = c0 * s; r.c1 = c1 * s; return r; }
    ZK_HD Fq2 inverse() const {
        Fq t = (c0.sqr() + c1.sqr()).inverse();
        Fq2 r; r.c0 = c0 * t; r.c1 = (c1 * t).neg();
        return r;
    }
};

} // namespace zk
