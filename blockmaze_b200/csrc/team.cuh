// Team point operations: the four warps of a 128-thread CTA work TOGETHER on 32 independent chains of point additions.
//
// Why.  The tails of an MSM (folding the pieces of a bucket, the running sums and trees of the bucket reduction; BDLO12's
// libff/algebra/scalar_multiplication/multiexp.tcc:244-278) are chains of DEPENDENT point additions with far too few chains to fill the
// GPU: what they cost is the latency of one addition, 14 field multiplications one after the other = 14 400 clocks (7.3 us) for a warp
// that has its SM sub-partition to itself (profiles/r02_notes.md).  Inside one warp nothing helps: a multiplication keeps the
// sub-partition's multiply pipe busy for 550 clocks whatever the instruction-level parallelism, which is why the quad-cooperative addition
// of round 2 (four LANES per addition) only reached 9 400 clocks.  The four sub-partitions of an SM have four multiply pipes, though: here
// lane l of EVERY warp of the CTA belongs to chain l, each warp does one of the (up to four) independent multiplications of a level of the
// addition's dependency graph on its own sub-partition, and the products cross through shared memory.  An addition is 4 levels (a doubling 3)
// instead of 14 (9) multiplications deep.  The operands stay replicated in the registers of all four warps, so control flow is identical in
// the four warps and per-lane special cases (infinity, doubling) are ordinary predication -- no warp-uniform schedules as the quad version
// needed.
//
// Protocol: every thread of the CTA calls every team function (they contain __syncthreads).  A level writes its products to one half of the
// exchange buffer, synchronises, reads; the next level uses the other half, so a writer can never overtake a reader of the level before last.
#pragma once
#include "ec.cuh"

namespace zk {

template <class F> struct Team {
    static constexpr int Q = (int)(sizeof(F) / 16);                 // 16-byte pieces of a field element
    static constexpr int BYTES = 2 * 4 * Q * 32 * 16;                // two halves x four products x 32 chains: Fq 8 KB, Fq2 16 KB
    uint4 *ex;
    int ph;                                                          // half in use by the current level
    int w, l;                                                        // warp = which product of a level; lane = chain
    int serial = 0;                                                  // debugging aid: every warp runs the one-thread operations (no exchange)
    __device__ __forceinline__ explicit Team(void *smem) : ex(reinterpret_cast<uint4 *>(smem)), ph(0), w((int)(threadIdx.x >> 5)), l((int)(threadIdx.x & 31)) {}
    // [half][product][piece][chain]: consecutive lanes 16 bytes apart, conflict-free 128-bit accesses
    __device__ __forceinline__ void put(int slot, const F &x) const {
        const uint32_t *d = reinterpret_cast<const uint32_t *>(&x);
        uint4 *p = ex + (size_t)((ph * 4 + slot) * Q) * 32 + l;
#pragma unroll
        for (int q = 0; q < Q; q++) p[q * 32] = make_uint4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
    }
    __device__ __forceinline__ F get(int slot) const {
        F x;
        uint32_t *d = reinterpret_cast<uint32_t *>(&x);
        const uint4 *p = ex + (size_t)((ph * 4 + slot) * Q) * 32 + l;
#pragma unroll
        for (int q = 0; q < Q; q++) { const uint4 v = p[q * 32]; d[4 * q] = v.x; d[4 * q + 1] = v.y; d[4 * q + 2] = v.z; d[4 * q + 3] = v.w; }
        return x;
    }
    __device__ __forceinline__ void next() { ph ^= 1; }
};

// a += o (add-2008-s, the formulas of XYZZ::add in ec.cuh, same result bit for bit).  Levels:
//   1: U1 = X1 ZZ2 | U2 = X2 ZZ1 | S1 = Y1 ZZZ2 | S2 = Y2 ZZZ1          P = U2 - U1, R = S2 - S1
//   2: PP = P^2    | RR = R^2    | ZZ1 ZZ2      | ZZZ1 ZZZ2
//   3: PPP = P PP  | Q = U1 PP   | ZZ3 = ZZ1 ZZ2 PP |  -                 X3 = RR - PPP - 2Q
//   4: R (Q - X3)  | S1 PPP      |  -           | ZZZ3 = ZZZ1 ZZZ2 PPP   Y3 = the difference of the first two
//
// Operands and result travel BY VALUE through the out-of-line functions: pointers to the caller's registers-turned-local-memory are what
// the miscompiled nested call of ec.cuh (see XYZZ::dbl_value) tripped over, and a first version of the reduction kernel on reference
// parameters gave run-to-run different sums with clean memcheck/racecheck reports.  An addition has four levels, so the exchange half in
// use is the same before and after; a doubling has three and flips it (the wrappers below keep the caller's copy in step).
template <class F> __device__ __noinline__ XYZZ<F> team_add_v(Team<F> tm, const XYZZ<F> a, const XYZZ<F> o) {
    if (tm.serial) { XYZZ<F> t = a; t.add(o); return t; }
    const bool o_inf = o.is_inf(), a_inf = a.is_inf();
    F m;
    switch (tm.w) {
        case 0: m = a.X * o.ZZ; break;
        case 1: m = o.X * a.ZZ; break;
        case 2: m = a.Y * o.ZZZ; break;
        default: m = o.Y * a.ZZZ; break;
    }
    tm.put(tm.w, m);
    __syncthreads();
    const F U1 = tm.get(0), S1 = tm.get(2);
    const F Pp = tm.get(1) - U1, R = tm.get(3) - S1;
    tm.next();
    switch (tm.w) {
        case 0: m = Pp.sqr(); break;
        case 1: m = R.sqr(); break;
        case 2: m = a.ZZ * o.ZZ; break;
        default: m = a.ZZZ * o.ZZZ; break;
    }
    if (tm.w < 2) tm.put(tm.w, m);
    __syncthreads();
    const F PP = tm.get(0), RR = tm.get(1);
    tm.next();
    F m3 = m;
    switch (tm.w) {
        case 0: m3 = Pp * PP; break;
        case 1: m3 = U1 * PP; break;
        case 2: m3 = m * PP; break;
        default: break;
    }
    if (tm.w < 3) tm.put(tm.w, m3);
    __syncthreads();
    const F PPP = tm.get(0), Q = tm.get(1);
    XYZZ<F> r;
    r.ZZ = tm.get(2);
    tm.next();
    r.X = RR - PPP - Q.dbl();
    switch (tm.w) {
        case 0: m3 = R * (Q - r.X); break;
        case 1: m3 = S1 * PPP; break;
        case 3: m3 = m * PPP; break;
        default: break;
    }
    if (tm.w != 2) tm.put(tm.w, m3);
    __syncthreads();
    r.Y = tm.get(0) - tm.get(1);
    r.ZZZ = tm.get(3);
    tm.next();
    if (!o_inf && !a_inf && Pp.is_zero()) {                          // same x: a doubling or a cancellation (rare; every warp redoes it alone)
        if (R.is_zero()) r = XYZZ<F>::dbl_value(a); else r = XYZZ<F>::inf();
    }
    return a_inf ? o : (o_inf ? a : r);
}
template <class F> __device__ __forceinline__ void team_add(Team<F> &tm, XYZZ<F> &a, const XYZZ<F> &o) { a = team_add_v(tm, a, o); }

// a = 2a (dbl-2008-s-1, the formulas of XYZZ::dbl).  Levels:
//   1: V = U^2 (U = 2Y) | XX = X^2                                     M = 3 XX
//   2: W = U V          | S = X V    | M^2     | ZZ3 = V ZZ             X3 = M^2 - 2S
//   3: M (S - X3)       | W Y        | ZZZ3 = W ZZZ                     Y3 = the difference of the first two
template <class F> __device__ __noinline__ XYZZ<F> team_dbl_v(Team<F> tm, const XYZZ<F> a) {
    if (tm.serial) return XYZZ<F>::dbl_value(a);
    const bool inf = a.is_inf();
    const F U = a.Y.dbl();
    F m = U;
    switch (tm.w) {
        case 0: m = U.sqr(); break;
        case 1: m = a.X.sqr(); break;
        default: break;
    }
    if (tm.w < 2) tm.put(tm.w, m);
    __syncthreads();
    const F V = tm.get(0), XX = tm.get(1);
    const F M = XX.dbl() + XX;
    tm.next();
    switch (tm.w) {
        case 0: m = U * V; break;
        case 1: m = a.X * V; break;
        case 2: m = M.sqr(); break;
        default: m = V * a.ZZ; break;
    }
    tm.put(tm.w, m);
    __syncthreads();
    const F W = tm.get(0), S = tm.get(1), MM = tm.get(2);
    XYZZ<F> r;
    r.ZZ = tm.get(3);
    tm.next();
    r.X = MM - S.dbl();
    switch (tm.w) {
        case 0: m = M * (S - r.X); break;
        case 1: m = W * a.Y; break;
        case 2: m = W * a.ZZZ; break;
        default: break;
    }
    if (tm.w < 3) tm.put(tm.w, m);
    __syncthreads();
    r.Y = tm.get(0) - tm.get(1);
    r.ZZZ = tm.get(2);
    return inf ? a : r;
}
template <class F> __device__ __forceinline__ void team_dbl(Team<F> &tm, XYZZ<F> &a) { a = team_dbl_v(tm, a); if (!tm.serial) tm.next(); }

// k * p per chain, k < 2^(top+1) for every chain of the CTA (top is CTA-uniform): double-and-add from the top bit; an addition is only
// executed at bit positions where some chain of the CTA has a one
template <class F> __device__ __forceinline__ XYZZ<F> team_mul_small(Team<F> &tm, const XYZZ<F> &p, uint32_t k, int top) {
    XYZZ<F> r = XYZZ<F>::inf();
    bool started = false;
    for (int b = top; b >= 0; b--) {
        if (started) team_dbl(tm, r);
        const bool bit = (k >> b) & 1u;
        if (__syncthreads_or(bit)) {
            const XYZZ<F> o = bit ? p : XYZZ<F>::inf();
            team_add(tm, r, o);
            started = true;
        }
    }
    return r;
}

// sum over the 32 chains of the CTA; `tree` = 32 XYZZ<F> of shared memory.  The total ends up in chain 0 (of every warp).
template <class F> __device__ __forceinline__ void team_tree_sum(Team<F> &tm, XYZZ<F> &a, XYZZ<F> *tree) {
    for (int d = 16; d > 0; d >>= 1) {
        if (tm.w == 0 && tm.l >= d && tm.l < 2 * d) tree[tm.l] = a;
        __syncthreads();
        const XYZZ<F> o = tm.l < d ? tree[tm.l + d] : XYZZ<F>::inf();
        team_add(tm, a, o);                                           // (its barriers order these reads before the next level's writes)
    }
}

} // namespace zk
