// Pippenger multi-scalar multiplication for alt_bn128 G1 / G2 on sm_100a.
//
// Replaces libff::multi_exp<.., BDLO12> and its 0/1 pre-filter (libff/algebra/scalar_multiplication/multiexp.tcc:165-282,
// 443-496) and libsnark's kc_multi_exp_with_mixed_addition (libsnark/knowledge_commitment/kc_multiexp.tcc:21-89).
// The result is the same group element; the schedule is GPU-shaped:
//   1. digits   : every scalar is cut into signed c-bit digits d_k in [-2^(c-1), 2^(c-1)].  Scalars equal to 0 are dropped;
//                 scalars equal to 1 (45 % of a BlockMaze witness) are spread round-robin over a separate range of "ones"
//                 buckets instead of all landing in bucket (0, 1).
//                 Two bucket layouts:
//                   windowed : digit (k, |d|) -> bucket k*2^(c-1) + |d| - 1, base i            (bases given per call)
//                   expanded : digit (k, |d|) -> bucket |d| - 1,             base k*n + i      (bases fixed in a proving key:
//                              2^(c*k) * P_i is precomputed once, so all windows share ONE bucket set and no Horner
//                              recombination is left -- fewer buckets to reduce, nothing to do on the host)
//   2. count / scan / scatter : a counting sort of (bucket -> base index | sign) built from global atomics.
//   3. accumulate : the sorted entry list is cut into equal ranges, one per thread of a single resident wave; a thread does
//                   XYZZ += affine mixed additions (8M+2S) and emits one piece per bucket it touches; the pieces of a bucket
//                   are then folded (small spans by one thread, oversized buckets by a CTA tree).
//   4. reduce  : per bucket region, sum_j (j+1)*B_j by segmented running sums + a shared-memory tree; ones buckets are summed.
//   5. the partial sums (a few dozen points) go back to the host.
#pragma once
#include <cuda_runtime.h>
#include "ec.cuh"

namespace zk {

struct MsmShape {
    int c;               // window bits
    int windows;         // ceil(255 / c)
    uint32_t nb;         // buckets per region = 2^(c-1)
    uint32_t ones;       // number of "ones" buckets
    uint32_t regions;    // bucket regions: 1 if expanded, else `windows`
    uint32_t total;      // regions*nb + ones
    uint32_t n;          // number of points
    int expanded;
};
static inline MsmShape msm_shape(uint32_t n, int c, uint32_t ones, int expanded) {
    MsmShape s; s.c = c; s.windows = (255 + c - 1) / c; s.nb = 1u << (c - 1); s.ones = ones; s.n = n; s.expanded = expanded;
    s.regions = expanded ? 1u : (uint32_t)s.windows; s.total = s.regions * s.nb + ones;
    return s;
}

__device__ __forceinline__ void ld_scalar(const uint32_t *p, uint32_t s[8]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

// Visit the non-zero signed digits of the canonical scalar s of point i.  f(bucket, entry)   entry = base index | sign << 31
template <class Fn>
__device__ __forceinline__ void msm_for_digits(const uint32_t s[8], const MsmShape &sh, uint32_t i, Fn f) {
    uint32_t orv = s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7];
    if (orv == 0) {
        if (s[0] == 0) return;
        if (s[0] == 1 && sh.ones) { f(sh.regions * sh.nb + (i % sh.ones), i); return; }
    }
    uint32_t carry = 0;
    const uint32_t mask = (1u << sh.c) - 1;
    for (int k = 0; k < sh.windows; k++) {
        const int bit = k * sh.c;
        const int w = bit >> 5, off = bit & 31;
        uint32_t v = 0;
        if (w < 8) {
            v = s[w] >> off;
            if (off + sh.c > 32 && w + 1 < 8) v |= s[w + 1] << (32 - off);
        }
        v = (v & mask) + carry;
        carry = 0;
        uint32_t neg = 0;
        if (v > sh.nb) { v = (1u << sh.c) - v; neg = 0x80000000u; carry = 1; }
        if (v != 0) {
            if (sh.expanded) f(v - 1, ((uint32_t)k * sh.n + i) | neg);
            else f((uint32_t)k * sh.nb + v - 1, i | neg);
        }
    }
}

// scalar source: scalars[(map ? map[i] : i + offset)], 8 words each, canonical (or Montgomery if `montgomery`)
struct ScalarSrc {
    const uint32_t *scalars;
    const uint32_t *map;
    uint32_t offset;
    int montgomery;
};
__device__ __forceinline__ void msm_load_scalar(const ScalarSrc &src, uint32_t i, uint32_t s[8]) {
    uint32_t j = src.map ? __ldg(src.map + i) : i + src.offset;
    ld_scalar(src.scalars + (size_t)j * 8, s);
    if (src.montgomery) {
        Fr x; for (int w = 0; w < 8; w++) x.v[w] = s[w];
        x = x.from_mont();
        for (int w = 0; w < 8; w++) s[w] = x.v[w];
    }
}

// skip[i] != 0 marks a base that is the point at infinity (its scalar is ignored)
static __global__ void msm_count_kernel(ScalarSrc src, const uint8_t *skip, MsmShape sh, uint32_t *counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sh.n) return;
    if (skip && skip[i]) return;
    uint32_t s[8];
    msm_load_scalar(src, i, s);
    msm_for_digits(s, sh, i, [&](uint32_t bucket, uint32_t) { atomicAdd(counts + bucket, 1u); });
}
static __global__ void msm_scatter_kernel(ScalarSrc src, const uint8_t *skip, MsmShape sh, const uint32_t *offsets, uint32_t *cursors,
                                          uint32_t *entries) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sh.n) return;
    if (skip && skip[i]) return;
    uint32_t s[8];
    msm_load_scalar(src, i, s);
    msm_for_digits(s, sh, i, [&](uint32_t bucket, uint32_t entry) {
        uint32_t pos = atomicAdd(cursors + bucket, 1u);
        entries[offsets[bucket] + pos] = entry;
    });
}

// The scatter with every atomic of a scalar in flight at once: window count and width are compile-time, so the digit loop unrolls, the
// up to W cursor increments are issued back to back (each returns the entry's ABSOLUTE position: the scan seeded the cursors with the
// bucket offsets) and the W stores follow.  msm_scatter_kernel above waits for one atomic's return per digit, W round trips to L2 in a row.
template <int C>
static __global__ void __launch_bounds__(256) msm_scatter_abs_kernel(ScalarSrc src, const uint8_t *skip, MsmShape sh, uint32_t *cursors,
                                                                     uint32_t *__restrict__ entries) {
    constexpr int W = (255 + C - 1) / C;
    constexpr uint32_t NONE = 0xffffffffu, MASK = (1u << C) - 1, NB = 1u << (C - 1);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sh.n) return;
    if (skip && skip[i]) return;
    uint32_t s[8];
    msm_load_scalar(src, i, s);
    if ((s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) {
        if (s[0] == 0) return;
        if (s[0] == 1 && sh.ones) { entries[atomicAdd(cursors + sh.regions * NB + (i % sh.ones), 1u)] = i; return; }
    }
    uint32_t at[W], en[W];
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < W; k++) {
        const int bit = k * C, w = bit >> 5, off = bit & 31;
        uint32_t v = 0;
        if (w < 8) {
            v = s[w] >> off;
            if (off + C > 32 && w + 1 < 8) v |= s[w + 1] << (32 - off);
        }
        v = (v & MASK) + carry;
        carry = 0;
        uint32_t neg = 0;
        if (v > NB) { v = (1u << C) - v; neg = 0x80000000u; carry = 1; }
        at[k] = NONE; en[k] = 0;
        if (v != 0) {
            if (sh.expanded) { at[k] = v - 1; en[k] = ((uint32_t)k * sh.n + i) | neg; }
            else { at[k] = (uint32_t)k * NB + v - 1; en[k] = i | neg; }
        }
    }
#pragma unroll
    for (int k = 0; k < W; k++) if (at[k] != NONE) at[k] = atomicAdd(cursors + at[k], 1u);
#pragma unroll
    for (int k = 0; k < W; k++) if (at[k] != NONE) entries[at[k]] = en[k];
}

// exclusive scan of `n` counts by one CTA of 32 warps; every warp owns a contiguous slice and reads it coalesced.  offsets[n] = total
// pad_log > 0 (affine rounds, below): every non-empty bucket's run is padded to a multiple of 2^pad_log entries, and offsets_shifted (if not
// null) receives offsets >> pad_log, the bucket boundaries after pad_log halving rounds.
// cursors != null (msm_scatter_abs_kernel follows): cursors = offsets, the counts are cleared for the next run and so are the two queues of
// oversized buckets -- no memset is left on the stream between the scalars and the accumulation.
static __global__ void __launch_bounds__(1024) msm_scan_kernel(uint32_t *counts, uint32_t *__restrict__ offsets, uint32_t n,
                                                               int pad_log = 0, uint32_t *__restrict__ offsets_shifted = nullptr,
                                                               uint32_t *__restrict__ cursors = nullptr, uint32_t *heavy_a = nullptr, uint32_t *heavy_b = nullptr) {
    const uint32_t pad = (1u << pad_log) - 1u;
    __shared__ uint32_t warp_off[33];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t per = (((n + 31) / 32) + 31) / 32 * 32;
    const uint32_t lo = min(warp * per, n), hi = min(lo + per, n);
    uint32_t sum = 0;
    for (uint32_t i = lo + lane; i < hi; i += 32) sum += (counts[i] + pad) & ~pad;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) warp_off[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = warp_off[lane], incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += t; }
        warp_off[lane] = incl - v;
        if (lane == 31) warp_off[32] = incl;
    }
    __syncthreads();
    uint32_t run = warp_off[warp];
    for (uint32_t base = lo; base < hi; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t v = i < hi ? (counts[i] + pad) & ~pad : 0;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, d); if ((int)lane >= d) incl += t; }
        if (i < hi) {
            offsets[i] = run + incl - v; if (offsets_shifted) offsets_shifted[i] = (run + incl - v) >> pad_log;
            if (cursors) { cursors[i] = run + incl - v; counts[i] = 0; }
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (threadIdx.x == 0) {
        offsets[n] = warp_off[32]; if (offsets_shifted) offsets_shifted[n] = warp_off[32] >> pad_log;
        if (heavy_a) heavy_a[0] = 0;
        if (heavy_b) heavy_b[0] = 0;
    }
}

template <class F> __device__ __forceinline__ Affine<F> ld_affine(const Affine<F> *p) {
    Affine<F> a;
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint32_t *d = reinterpret_cast<uint32_t *>(&a);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(Affine<F>) / 16); i++) {
        uint4 v = __ldg(q + i);
        d[4 * i] = v.x; d[4 * i + 1] = v.y; d[4 * i + 2] = v.z; d[4 * i + 3] = v.w;
    }
    return a;
}
__device__ __forceinline__ Fq ldg_fq(const Fq *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    Fq r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ Fq ld_fq_plain(const Fq *p) {               // written earlier by this very thread: no read-only path
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = q[0], b = q[1];
    Fq r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fq(Fq *p, const Fq &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ void st_affine(Affine<Fq> *p, const Affine<Fq> &a) { st_fq(&p->x, a.x); st_fq(&p->y, a.y); }
template <class F> __device__ __forceinline__ void st_xyzz(XYZZ<F> *p, const XYZZ<F> &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    const uint32_t *d = reinterpret_cast<const uint32_t *>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 16); i++) q[i] = make_uint4(d[4 * i], d[4 * i + 1], d[4 * i + 2], d[4 * i + 3]);
}
template <class F> __device__ __forceinline__ XYZZ<F> ld_xyzz(const XYZZ<F> *p) {
    XYZZ<F> v;
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint32_t *d = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 16); i++) { uint4 x = q[i]; d[4 * i] = x.x; d[4 * i + 1] = x.y; d[4 * i + 2] = x.z; d[4 * i + 3] = x.w; }
    return v;
}

// ---- fixed-base expansion (proving-key load time) --------------------------------------------------------------------------
// out[k*n + i] = 2^(c*k) * in[i] in affine form, k < windows.  One thread per point.
template <class F>
static __global__ void __launch_bounds__(128) msm_expand_bases_kernel(const Affine<F> *__restrict__ in, uint32_t n, int c, int windows,
                                                                      Affine<F> *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Affine<F> p = in[i];
    out[i] = p;
    XYZZ<F> cur = XYZZ<F>::from_affine(p);
    for (int k = 1; k < windows; k++) {
        for (int j = 0; j < c; j++) cur = cur.dbl();
        out[(size_t)k * n + i] = cur.to_affine();
    }
}

// ---- bucket accumulation, load-balanced ------------------------------------------------------------------------------------
// Bucket sizes are wildly uneven in practice (the top window of a 254-bit scalar has only 2-3 live bits, so a handful of
// buckets hold n/4 points each; witness scalars repeat), and the kernel is bound by the integer-multiply pipe, so what matters
// is that every resident thread gets the same number of mixed additions.  The SORTED ENTRY LIST -- not the bucket list -- is
// therefore cut into equal ranges of L = ceil(entries / T) entries, T = the threads of exactly one resident wave:
//   accumulate : thread t sums entries [t*L, (t+1)*L) (XYZZ += affine); whenever the range crosses into the next bucket it
//                stores the finished piece and starts over.  The piece of bucket b made by thread t lives in slot t + b
//                (both only grow along the list, so slots never collide).
//   fold       : bucket b owns the pieces of threads off[b]/L .. (off[b+1]-1)/L; they are summed into the first slot -- by
//                one thread when the span is small, by a whole CTA (strided sums + shared-memory tree) for the few oversized
//                buckets, which are queued in heavy[]
constexpr uint32_t MSM_MIN_RANGE = 16;
constexpr uint32_t MSM_FOLD_SMALL = 16;
constexpr uint32_t MSM_HEAVY_MAX = 1024;

__host__ __device__ __forceinline__ uint32_t msm_range_len(uint32_t total_entries, uint32_t threads) {
    const uint32_t L = (total_entries + threads - 1) / threads;
    return L < MSM_MIN_RANGE ? MSM_MIN_RANGE : L;
}

#ifndef ZK_ACC_MINBLOCKS
#define ZK_ACC_MINBLOCKS 1
#endif
#ifndef ZK_ACC_INLINE_MUL
#define ZK_ACC_INLINE_MUL 1      // G1: also inline the ten field multiplications of the mixed addition (30 KB loop body, 118 registers;
                                 // measured 10 % faster than calling the shared copy: no argument shuffling on the multiply pipe)
#endif
#ifndef ZK_ACC_INLINE_ADD
#define ZK_ACC_INLINE_ADD 1      // G1: keep the accumulator in registers across the range loop (measured 5 % faster; 108 registers)
#endif
template <class F>
static __global__ void __launch_bounds__(128, ZK_ACC_MINBLOCKS) msm_accumulate_kernel(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ offsets,
                                                                    const uint32_t *__restrict__ entries, uint32_t total_buckets, uint32_t threads,
                                                                    XYZZ<F> *__restrict__ partial) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = __ldg(offsets + total_buckets), L = msm_range_len(total, threads);
    const uint32_t e0 = t * L, e1 = min(e0 + L, total);
    if (t >= threads || e0 >= total) return;
    // bucket of entry e0: the last b with offsets[b] <= e0 (among equal offsets that is the non-empty one)
    uint32_t lo = 0, hi = total_buckets;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(offsets + mid) <= e0) lo = mid; else hi = mid; }
    uint32_t b = lo, next = __ldg(offsets + b + 1);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t e = e0; e < e1; e++) {
        if (e == next) {
            st_xyzz(partial + t + b, acc);
            acc = XYZZ<F>::inf();
            b++; next = __ldg(offsets + b + 1);
            if (next <= e) {                       // a run of empty buckets (sparse witness MSMs): search, do not walk
                lo = b; hi = total_buckets;
                while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(offsets + mid) <= e) lo = mid; else hi = mid; }
                b = lo; next = __ldg(offsets + b + 1);
            }
        }
        const uint32_t ent = entries ? __ldg(entries + e) : e;      // entries == nullptr: `bases` is already the bucket-sorted point list
        Affine<F> p = ld_affine(bases + (ent & 0x7fffffffu));
        if (ent & 0x80000000u) p.y = p.y.neg();
#if ZK_ACC_INLINE_MUL
        if constexpr (sizeof(F) == 32) acc.add_affine_with(p, [](const F &x, const F &y) { return F::mul_impl(x, y); }, [](const F &x) { return F::sqr_impl(x); },
                                                            [](const F &a, const F &b, const F &c, const F &d) { return F::mul2_impl(a, b, c, d); });
        else acc.add_affine(p);
#elif ZK_ACC_INLINE_ADD
        if (sizeof(F) == 32) acc.add_affine_inl(p); else acc.add_affine(p);
#else
        acc.add_affine(p);
#endif
    }
    st_xyzz(partial + t + b, acc);
}
// ---- batched-affine halving rounds (dense fixed-base MSM: the H query) -------------------------------------------------------------
// An XYZZ += affine addition is 10 multiplications; affine + affine is 5M + 1S + one inversion, and inversions batch (Montgomery's
// trick: 3 more multiplications per pair, ONE inversion per batch) -- 6 multiplications per addition if the inversion is cheap and the
// batch long.  On a SIMT machine an inversion shared by a warp costs what 32 private ones cost, so each THREAD batches its own pairs,
// and with ~55 pairs per thread only the safegcd inversion (ff.cuh inverse_gcd: ~1/25 of a Fermat inverse on the multiply pipe) pays.
// Layout: the digit sort pads every bucket's run to a multiple of 2^R entries (R rounds; pads are the null entry 0xffffffff = infinity),
// so a round simply adds items 2i and 2i+1 of the flat list -- partners are always in the same bucket, no segment logic -- and after R
// rounds bucket b owns items offsets[b] >> R .. offsets[b+1] >> R of the last list, which the XYZZ accumulate kernel then finishes.
// Thread t takes pairs t, t + T, t + 2T, ... (coalesced): pass 1 multiplies the denominators up and stores the running products, one
// inversion, pass 2 walks back.  P + P (equal points) uses the tangent slope, P + (-P) and infinities give their result directly.
constexpr uint32_t MSM_NULL_ENTRY = 0xffffffffu;
// One pair of a halving round: where its two points live and their coordinates as far as they have been fetched.
struct AffPair {
    uint32_t ea, eb;            // FIRST round: sorted entries (base index | sign << 31, or MSM_NULL_ENTRY); later rounds: item indices
    Fq ax, bx, ay, by;
};
template <bool FIRST> __device__ __forceinline__ void aff_fetch_idx(const uint32_t *__restrict__ entries, uint32_t p, AffPair &q) {
    if (FIRST) { const uint2 e = __ldg(reinterpret_cast<const uint2 *>(entries) + p); q.ea = e.x; q.eb = e.y; }
    else { q.ea = 2 * p; q.eb = 2 * p + 1; }
}
template <bool FIRST> __device__ __forceinline__ void aff_fetch_x(const Affine<Fq> *__restrict__ pts, AffPair &q) {
    q.ax = (FIRST && q.ea == MSM_NULL_ENTRY) ? Fq::zero() : ldg_fq(&pts[q.ea & 0x7fffffffu].x);
    q.bx = (FIRST && q.eb == MSM_NULL_ENTRY) ? Fq::zero() : ldg_fq(&pts[q.eb & 0x7fffffffu].x);
}
template <bool FIRST> __device__ __forceinline__ void aff_fetch_y(const Affine<Fq> *__restrict__ pts, AffPair &q) {
    q.ay = (FIRST && q.ea == MSM_NULL_ENTRY) ? Fq::zero() : ldg_fq(&pts[q.ea & 0x7fffffffu].y);
    q.by = (FIRST && q.eb == MSM_NULL_ENTRY) ? Fq::zero() : ldg_fq(&pts[q.eb & 0x7fffffffu].y);
    if (FIRST) { if (q.ea != MSM_NULL_ENTRY && (q.ea & 0x80000000u)) q.ay = q.ay.neg(); if (q.eb != MSM_NULL_ENTRY && (q.eb & 0x80000000u)) q.by = q.by.neg(); }
}
// can the pair be classified from the x coordinates alone?  (generic chord: distinct x, neither point at infinity)
template <bool FIRST> __device__ __forceinline__ bool aff_is_generic(const AffPair &q) {
    if (FIRST) return q.ea != MSM_NULL_ENTRY && q.eb != MSM_NULL_ENTRY && q.ax != q.bx;
    return !q.ax.is_zero() && !q.bx.is_zero() && q.ax != q.bx;            // later rounds: infinity is (0, 0), so x != 0 rules it out
}
// kind of pair and its denominator; y coordinates must be present.  0 = chord, 1 = tangent (P + P), 2 = result a, 3 = result b, 4 = infinity
template <bool FIRST> __device__ __forceinline__ int aff_classify(const AffPair &q, Fq &den) {
    const bool ia = FIRST ? q.ea == MSM_NULL_ENTRY : (q.ax.is_zero() && q.ay.is_zero());
    const bool ib = FIRST ? q.eb == MSM_NULL_ENTRY : (q.bx.is_zero() && q.by.is_zero());
    den = Fq::one();
    if (ia && ib) return 4;
    if (ib) return 2;
    if (ia) return 3;
    if (q.ax != q.bx) { den = q.bx - q.ax; return 0; }
    if (q.ay == q.by && !q.ay.is_zero()) { den = q.ay.dbl(); return 1; }
    return 4;                                                              // P + (-P)
}
constexpr int AFF_UNROLL = 4;          // pairs whose loads are in flight together (the rounds are gather-latency bound otherwise)
template <bool FIRST>
static __global__ void __launch_bounds__(128) msm_affine_round_kernel(const Affine<Fq> *__restrict__ pts, const uint32_t *__restrict__ entries,
                                                                      const uint32_t *__restrict__ offsets, uint32_t total_buckets, int shift_in,
                                                                      Affine<Fq> *__restrict__ out, Fq *__restrict__ scratch, uint32_t threads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_pairs = (__ldg(offsets + total_buckets) >> shift_in) >> 1;
    if (t >= threads || t >= n_pairs) return;
    const uint32_t mine = (n_pairs - 1 - t) / threads + 1;                 // pairs t, t + T, ... of this thread
    // ---- pass 1: running product of the denominators ----
    Fq c = Fq::one();
    for (uint32_t k0 = 0; k0 < mine; k0 += AFF_UNROLL) {
        AffPair q[AFF_UNROLL];
#pragma unroll
        for (int u = 0; u < AFF_UNROLL; u++) if (k0 + u < mine) aff_fetch_idx<FIRST>(entries, t + (k0 + u) * threads, q[u]);
#pragma unroll
        for (int u = 0; u < AFF_UNROLL; u++) if (k0 + u < mine) aff_fetch_x<FIRST>(pts, q[u]);
#pragma unroll
        for (int u = 0; u < AFF_UNROLL; u++) {
            if (k0 + u >= mine) break;
            Fq den;
            if (aff_is_generic<FIRST>(q[u])) den = q[u].bx - q[u].ax;
            else { aff_fetch_y<FIRST>(pts, q[u]); aff_classify<FIRST>(q[u], den); }
            c = Fq::mul_impl(c, den);
            st_fq(scratch + t + (k0 + u) * threads, c);
        }
    }
    Fq inv = c.inverse_gcd();
    // ---- pass 2: walk back, one slope per pair ----
    for (uint32_t done = 0; done < mine; done += AFF_UNROLL / 2) {
        AffPair q[AFF_UNROLL / 2]; Fq prev[AFF_UNROLL / 2];
#pragma unroll
        for (int u = 0; u < AFF_UNROLL / 2; u++) if (done + u < mine) aff_fetch_idx<FIRST>(entries, t + (mine - 1 - done - u) * threads, q[u]);
#pragma unroll
        for (int u = 0; u < AFF_UNROLL / 2; u++) if (done + u < mine) {
            const uint32_t k = mine - 1 - done - u;
            aff_fetch_x<FIRST>(pts, q[u]); aff_fetch_y<FIRST>(pts, q[u]);
            if (k > 0) prev[u] = ld_fq_plain(scratch + t + (k - 1) * threads);
        }
#pragma unroll
        for (int u = 0; u < AFF_UNROLL / 2; u++) {
            if (done + u >= mine) break;
            const uint32_t k = mine - 1 - done - u;
            Fq den;
            const int kind = aff_classify<FIRST>(q[u], den);
            const Fq dinv = k > 0 ? Fq::mul_impl(inv, prev[u]) : inv;                    // inv = 1 / (product up to and including this pair)
            inv = Fq::mul_impl(inv, den);
            Affine<Fq> r;
            if (kind <= 1) {
                Fq num;
                if (kind == 0) num = q[u].by - q[u].ay; else { const Fq xx = q[u].ax.sqr(); num = xx.dbl() + xx; }
                const Fq lam = Fq::mul_impl(num, dinv);
                r.x = Fq::sqr_impl(lam) - q[u].ax - q[u].bx;
                r.y = Fq::mul_impl(lam, q[u].ax - r.x) - q[u].ay;
            } else if (kind == 2) { r.x = q[u].ax; r.y = q[u].ay; }
            else if (kind == 3) { r.x = q[u].bx; r.y = q[u].by; }
            else r = Affine<Fq>::inf();
            st_affine(out + t + k * threads, r);
        }
    }
}

// slots of bucket b: first slot and number of pieces (0 for an empty bucket)
__device__ __forceinline__ uint32_t msm_bucket_span(const uint32_t *__restrict__ offsets, uint32_t L, uint32_t b, uint32_t &slot0) {
    const uint32_t o0 = __ldg(offsets + b), o1 = __ldg(offsets + b + 1);
    if (o1 == o0) return 0;
    const uint32_t t0 = o0 / L;
    slot0 = t0 + b;
    return (o1 - 1) / L - t0 + 1;
}
// one thread per bucket: fold short spans in place, queue long ones
template <class F>
static __global__ void __launch_bounds__(128) msm_fold_small_kernel(XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ offsets,
                                                                    uint32_t total_buckets, uint32_t threads, uint32_t *__restrict__ heavy /* [0] = count */) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_buckets) return;
    const uint32_t L = msm_range_len(__ldg(offsets + total_buckets), threads);
    uint32_t t0 = 0;
    const uint32_t span = msm_bucket_span(offsets, L, b, t0);
    if (span <= 1) return;
    if (span > MSM_FOLD_SMALL) {
        const uint32_t slot = atomicAdd(heavy, 1u);
        if (slot < MSM_HEAVY_MAX) { heavy[1 + slot] = b; return; }      // (overflow: fall through and fold serially -- still correct)
    }
    XYZZ<F> a = ld_xyzz(partial + t0);
    for (uint32_t k = 1; k < span; k++) a.add(ld_xyzz(partial + t0 + k));
    st_xyzz(partial + t0, a);
}
constexpr int MSM_HEAVY_THREADS = 128;
template <class F>
static __global__ void __launch_bounds__(MSM_HEAVY_THREADS) msm_fold_heavy_kernel(XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ offsets,
                                                                                  uint32_t total_buckets, uint32_t threads, const uint32_t *__restrict__ heavy) {
    extern __shared__ uint32_t fold_sm[];
    XYZZ<F> *sm = reinterpret_cast<XYZZ<F> *>(fold_sm);
    const uint32_t cnt = min(heavy[0], MSM_HEAVY_MAX);
    const uint32_t L = msm_range_len(__ldg(offsets + total_buckets), threads);
    for (uint32_t h = blockIdx.x; h < cnt; h += gridDim.x) {
        uint32_t t0 = 0;
        const uint32_t span = msm_bucket_span(offsets, L, heavy[1 + h], t0);
        XYZZ<F> a = XYZZ<F>::inf();
        for (uint32_t k = threadIdx.x; k < span; k += MSM_HEAVY_THREADS) a.add(ld_xyzz(partial + t0 + k));
        sm[threadIdx.x] = a;
        __syncthreads();
        for (int d = MSM_HEAVY_THREADS / 2; d > 0; d >>= 1) {
            if ((int)threadIdx.x < d) { XYZZ<F> x = sm[threadIdx.x]; x.add(sm[threadIdx.x + d]); sm[threadIdx.x] = x; }
            __syncthreads();
        }
        if (threadIdx.x == 0) st_xyzz(partial + t0, sm[0]);
        __syncthreads();
    }
}
template <class F> __device__ __forceinline__ XYZZ<F> msm_bucket(const XYZZ<F> *partial, const uint32_t *offsets, uint32_t L, uint32_t b) {
    uint32_t t0 = 0;
    if (msm_bucket_span(offsets, L, b, t0) == 0) return XYZZ<F>::inf();
    return ld_xyzz(partial + t0);
}

// Per bucket region (blockIdx.y) and CTA (blockIdx.x): every thread owns `seg` consecutive buckets [lo, lo+seg) and computes
//   S = sum_j B_j   and   T = sum_j (j - lo + 1) * B_j     by the running-sum trick, then contributes  T + lo * S.
// The CTA's contributions are tree-summed in shared memory; one XYZZ per (region, blockIdx.x) is written.
// For blockIdx.y == regions the "ones" buckets are summed with weight 1.
constexpr int MSM_RED_THREADS = 128;
template <class F>
static __global__ void __launch_bounds__(MSM_RED_THREADS) msm_reduce_kernel(const XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ offsets,
                                                                            uint32_t threads, MsmShape sh, uint32_t seg, uint32_t seg_weighted,
                                                                            uint32_t blocks_per_region, XYZZ<F> *__restrict__ out) {
    extern __shared__ uint32_t red_sm[];
    XYZZ<F> *sm = reinterpret_cast<XYZZ<F> *>(red_sm);
    const uint32_t w = blockIdx.y;
    const bool ones = (w == sh.regions);
    const uint32_t count = ones ? sh.ones : sh.nb;
    const uint32_t base = w * sh.nb;                              // the ones region starts at regions*nb as well
    if (!ones) seg = seg_weighted;                                // (small weighted regions take one bucket per thread: shorter chains)
    const uint32_t lo = (blockIdx.x * MSM_RED_THREADS + threadIdx.x) * seg;
    const uint32_t L = msm_range_len(__ldg(offsets + sh.total), threads);
    XYZZ<F> S = XYZZ<F>::inf(), T = XYZZ<F>::inf();
    if (lo < count) {
        const uint32_t hi = min(lo + seg, count);
        for (uint32_t j = hi; j-- > lo;) {
            S.add(msm_bucket(partial, offsets, L, base + j));
            if (!ones) T.add(S);
        }
        if (ones) T = S;
        else if (lo && !S.is_inf()) T.add(S.mul_small(lo));
    }
    sm[threadIdx.x] = T;
    __syncthreads();
    for (int d = MSM_RED_THREADS / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) { XYZZ<F> a = sm[threadIdx.x]; a.add(sm[threadIdx.x + d]); sm[threadIdx.x] = a; }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_xyzz(out + (size_t)w * blocks_per_region + blockIdx.x, sm[0]);
}

} // namespace zk
