// Pippenger multi-scalar multiplication for alt_bn128 G1 / G2 on sm_100a.
//
// Replaces libff::multi_exp<.., BDLO12> and its 0/1 pre-filter (libff/algebra/scalar_multiplication/multiexp.tcc:165-282,
// 443-496) and libsnark's kc_multi_exp_with_mixed_addition (libsnark/knowledge_commitment/kc_multiexp.tcc:21-89).
// The result is the same group element; the schedule is GPU-shaped:
//   1. digits   : every scalar is cut into signed c-bit digits d_k in [-2^(c-1), 2^(c-1)].  Digit (k, |d|) selects bucket
//                 k*2^(c-1) + |d| - 1.  Scalars equal to 0 are dropped; scalars equal to 1 (45 % of a BlockMaze witness) are
//                 spread round-robin over a separate range of "ones" buckets instead of all landing in bucket (0,1).
//   2. count / scan / scatter : a counting sort of (bucket -> point index | sign) built from global atomics.
//   3. accumulate : buckets are cut into tasks of <= 48 entries; one thread per task does XYZZ += affine mixed additions
//                   (8M+2S) and a few segmented pairwise passes fold the tasks of oversized buckets.
//   4. reduce  : per window, sum_j (j+1)*B_j by segmented running sums + a shared-memory tree; ones buckets are summed.
//   5. the per-window partial sums (a few dozen points) go back to the host, which does the final Horner combination.
#pragma once
#include <cuda_runtime.h>
#include "ec.cuh"

namespace zk {

struct MsmShape {
    int c;               // window bits
    int windows;         // ceil(255 / c)
    uint32_t nb;         // buckets per window = 2^(c-1)
    uint32_t ones;       // number of "ones" buckets
    uint32_t total;      // windows*nb + ones
};
static inline MsmShape msm_shape(int c, uint32_t ones) {
    MsmShape s; s.c = c; s.windows = (255 + c - 1) / c; s.nb = 1u << (c - 1); s.ones = ones; s.total = s.windows * s.nb + ones; return s;
}

__device__ __forceinline__ void ld_scalar(const uint32_t *p, uint32_t s[8]) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    s[0] = a.x; s[1] = a.y; s[2] = a.z; s[3] = a.w; s[4] = b.x; s[5] = b.y; s[6] = b.z; s[7] = b.w;
}

// Visit the non-zero signed digits of the canonical scalar s.  f(bucket, negative)
template <class Fn>
__device__ __forceinline__ void msm_for_digits(const uint32_t s[8], const MsmShape &sh, uint32_t point_idx, Fn f) {
    uint32_t orv = s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7];
    if (orv == 0) {
        if (s[0] == 0) return;
        if (s[0] == 1 && sh.ones) { f(sh.windows * sh.nb + (point_idx % sh.ones), false); return; }
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
        bool neg = false;
        if (v > sh.nb) { v = (1u << sh.c) - v; neg = true; carry = 1; }
        if (v != 0) f((uint32_t)k * sh.nb + v - 1, neg);
    }
}

// scalar source: scalars[(map ? map[i] : i + offset)], 8 words each, canonical (or Montgomery if from_mont)
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
static __global__ void msm_count_kernel(ScalarSrc src, const uint8_t *skip, uint32_t n, MsmShape sh, uint32_t *counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (skip && skip[i]) return;
    uint32_t s[8];
    msm_load_scalar(src, i, s);
    msm_for_digits(s, sh, i, [&](uint32_t bucket, bool) { atomicAdd(counts + bucket, 1u); });
}
static __global__ void msm_scatter_kernel(ScalarSrc src, const uint8_t *skip, uint32_t n, MsmShape sh, const uint32_t *offsets, uint32_t *cursors,
                                   uint32_t *entries) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (skip && skip[i]) return;
    uint32_t s[8];
    msm_load_scalar(src, i, s);
    msm_for_digits(s, sh, i, [&](uint32_t bucket, bool neg) {
        uint32_t pos = atomicAdd(cursors + bucket, 1u);
        entries[offsets[bucket] + pos] = i | (neg ? 0x80000000u : 0u);
    });
}

// exclusive scan of `n` counts by one CTA of 1024 threads (n up to a few million): offsets[n] = total
static __global__ void __launch_bounds__(1024) msm_scan_kernel(const uint32_t *counts, uint32_t *offsets, uint32_t n) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, n);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += counts[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        uint32_t v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (uint32_t i = lo; i < hi; i++) { offsets[i] = run; run += counts[i]; }
    if (threadIdx.x == 1023) offsets[n] = part[1023];
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

// ---- bucket accumulation, load-balanced ------------------------------------------------------------------------------
// Bucket sizes are wildly uneven in practice (the top window of a 254-bit scalar has only 2-3 live bits, so a handful of
// buckets hold n/4 points each; witness scalars repeat), so buckets are cut into TASKS of at most MSM_TASK entries:
//   task_count : tasks_b = ceil(count_b / MSM_TASK), scanned into task_off[]
//   accumulate : one thread per task sums its <= MSM_TASK points (XYZZ += affine) into partial[task]
//   combine    : log2(max tasks per bucket) segmented pairwise passes fold partial[] so that partial[task_off[b]] = bucket b
constexpr uint32_t MSM_TASK = 48;

static __global__ void msm_task_count_kernel(const uint32_t *__restrict__ offsets, uint32_t total_buckets, uint32_t *__restrict__ task_counts,
                                             uint32_t *__restrict__ max_tasks) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_buckets) return;
    const uint32_t cnt = offsets[b + 1] - offsets[b];
    const uint32_t t = (cnt + MSM_TASK - 1) / MSM_TASK;
    task_counts[b] = t;
    if (t > 1) atomicMax(max_tasks, t);
}

template <class F>
static __global__ void __launch_bounds__(128) msm_accumulate_kernel(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ offsets,
                                                             const uint32_t *__restrict__ entries, const uint32_t *__restrict__ task_off,
                                                             uint32_t total_buckets, XYZZ<F> *__restrict__ partial,
                                                             uint32_t *__restrict__ task_rank, uint32_t *__restrict__ task_span) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= task_off[total_buckets]) return;
    // bucket of task t: last b with task_off[b] <= t
    uint32_t lo = 0, hi = total_buckets;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(task_off + mid) <= t) lo = mid; else hi = mid; }
    const uint32_t b = lo, rank = t - __ldg(task_off + b);
    const uint32_t e0 = offsets[b] + rank * MSM_TASK, e1 = min(e0 + MSM_TASK, offsets[b + 1]);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t e = e0; e < e1; e++) {
        const uint32_t ent = __ldg(entries + e);
        Affine<F> p = ld_affine(bases + (ent & 0x7fffffffu));
        if (ent & 0x80000000u) p.y = p.y.neg();
        acc.add_affine(p);
    }
    st_xyzz(partial + t, acc);
    task_rank[t] = rank;
    task_span[t] = __ldg(task_off + b + 1) - __ldg(task_off + b);
}
// pass with stride s: partial[t] += partial[t + s] when rank % 2s == 0 and rank + s < span
template <class F>
static __global__ void __launch_bounds__(128) msm_combine_kernel(XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ task_rank,
                                                          const uint32_t *__restrict__ task_span, const uint32_t *__restrict__ task_off,
                                                          uint32_t total_buckets, const uint32_t *__restrict__ max_tasks, uint32_t s) {
    if (s >= *max_tasks) return;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= task_off[total_buckets]) return;
    const uint32_t rank = task_rank[t];
    if ((rank & (2 * s - 1)) != 0 || rank + s >= task_span[t]) return;
    XYZZ<F> a = ld_xyzz(partial + t);
    a.add(ld_xyzz(partial + t + s));
    st_xyzz(partial + t, a);
}
template <class F> __device__ __forceinline__ XYZZ<F> msm_bucket(const XYZZ<F> *partial, const uint32_t *task_off, uint32_t b) {
    const uint32_t t0 = __ldg(task_off + b);
    if (__ldg(task_off + b + 1) == t0) return XYZZ<F>::inf();
    return ld_xyzz(partial + t0);
}

// Per window w (blockIdx.y) and segment block (blockIdx.x): every thread owns `seg` consecutive buckets [lo, lo+seg) and computes
//   S = sum_j B_j   and   T = sum_j (j - lo + 1) * B_j     by the running-sum trick, then contributes  T + lo * S.
// The CTA's contributions are tree-summed in shared memory; one XYZZ per (window, blockIdx.x) is written.
// For blockIdx.y == windows the "ones" buckets are summed with weight 1.
constexpr int MSM_RED_THREADS = 64;
template <class F>
static __global__ void __launch_bounds__(MSM_RED_THREADS) msm_reduce_kernel(const XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ task_off, MsmShape sh,
                                                                     uint32_t seg, uint32_t blocks_per_window, XYZZ<F> *__restrict__ out) {
    extern __shared__ uint32_t red_sm[];
    XYZZ<F> *sm = reinterpret_cast<XYZZ<F> *>(red_sm);
    const uint32_t w = blockIdx.y;
    const bool ones = (w == (uint32_t)sh.windows);
    const uint32_t count = ones ? sh.ones : sh.nb;
    const uint32_t base = w * sh.nb;                              // ones region starts at windows*nb as well
    const uint32_t lo = (blockIdx.x * MSM_RED_THREADS + threadIdx.x) * seg;
    XYZZ<F> S = XYZZ<F>::inf(), T = XYZZ<F>::inf();
    if (lo < count) {
        const uint32_t hi = min(lo + seg, count);
        for (uint32_t j = hi; j-- > lo;) {
            S.add(msm_bucket(partial, task_off, base + j));
            if (!ones) T.add(S);
        }
        if (ones) T = S;
        else if (lo) T.add(S.mul_small(lo));
    }
    sm[threadIdx.x] = T;
    __syncthreads();
    for (int d = MSM_RED_THREADS / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) { XYZZ<F> a = sm[threadIdx.x]; a.add(sm[threadIdx.x + d]); sm[threadIdx.x] = a; }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_xyzz(out + (size_t)w * blocks_per_window + blockIdx.x, sm[0]);
}

} // namespace zk
