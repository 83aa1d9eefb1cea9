// The tails of an MSM on team point operations (team.cuh): fold of the bucket pieces, bucket reduction and the final sum.
// Same inputs, same group elements out as msm_fold_small/heavy_kernel and msm_reduce_kernel of msm.cuh (kept: ZKB200_TEAM=0 selects them);
// a CTA of 128 threads carries 32 chains, and the reduction finishes ON THE DEVICE: the last CTA of a bucket region adds the region's CTA
// sums, so one point per region goes back to the host instead of one per CTA.
// Reference: libff/algebra/scalar_multiplication/multiexp.tcc:244-278 (running sums over the buckets of BDLO12).
#pragma once
#include "msm.cuh"
#include "team.cuh"

namespace zk {

constexpr int TEAM_CHAINS = 32;
constexpr uint32_t MSM_TEAM_FOLD_MAX_BUCKETS = 8192;       // above this the piece fold is throughput-bound: one thread per bucket
template <class F> constexpr size_t team_smem_bytes() { return (size_t)Team<F>::BYTES + TEAM_CHAINS * sizeof(XYZZ<F>); }

// chain = bucket: the pieces of a bucket are added up in place; oversized buckets are queued for the heavy kernel
template <class F>
static __global__ void __launch_bounds__(128) msm_fold_small_team_kernel(XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ offsets,
                                                                         uint32_t total_buckets, uint32_t threads, uint32_t *__restrict__ heavy) {
    extern __shared__ uint4 team_sm[];
    __shared__ uint32_t queued[TEAM_CHAINS];
    Team<F> tm(team_sm);
    const uint32_t b = blockIdx.x * TEAM_CHAINS + tm.l;
    const uint32_t L = msm_range_len(__ldg(offsets + total_buckets), threads);
    uint32_t t0 = 0;
    uint32_t span = b < total_buckets ? msm_bucket_span(offsets, L, b, t0) : 0;
    if (tm.w == 0) {
        uint32_t q = 0;
        if (span > MSM_FOLD_SMALL) {
            const uint32_t slot = atomicAdd(heavy, 1u);
            if (slot < MSM_HEAVY_MAX) { heavy[1 + slot] = b; q = 1; }        // (overflow: folded here, chain by chain -- still correct)
        }
        queued[tm.l] = q;
    }
    __syncthreads();
    if (queued[tm.l]) span = 0;
    XYZZ<F> a = span > 1 ? ld_xyzz(partial + t0) : XYZZ<F>::inf();
    for (uint32_t k = 1; __syncthreads_or(k < span); k++) {
        const XYZZ<F> o = k < span ? ld_xyzz(partial + t0 + k) : XYZZ<F>::inf();
        team_add(tm, a, o);
    }
    if (tm.w == 0 && span > 1) st_xyzz(partial + t0, a);
}

// CTA = one oversized bucket at a time: chain l adds pieces l, l + 32, ...; tree over the chains
template <class F>
static __global__ void __launch_bounds__(128) msm_fold_heavy_team_kernel(XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ offsets,
                                                                         uint32_t total_buckets, uint32_t threads, const uint32_t *__restrict__ heavy) {
    extern __shared__ uint4 team_sm[];
    Team<F> tm(team_sm);
    XYZZ<F> *tree = reinterpret_cast<XYZZ<F> *>(reinterpret_cast<char *>(team_sm) + Team<F>::BYTES);
    const uint32_t cnt = min(heavy[0], MSM_HEAVY_MAX);
    const uint32_t L = msm_range_len(__ldg(offsets + total_buckets), threads);
    for (uint32_t h = blockIdx.x; h < cnt; h += gridDim.x) {
        uint32_t t0 = 0;
        const uint32_t span = msm_bucket_span(offsets, L, heavy[1 + h], t0);
        XYZZ<F> a = XYZZ<F>::inf();
        for (uint32_t k0 = 0; k0 < span; k0 += TEAM_CHAINS) {
            const XYZZ<F> o = k0 + tm.l < span ? ld_xyzz(partial + t0 + k0 + tm.l) : XYZZ<F>::inf();
            team_add(tm, a, o);
        }
        team_tree_sum(tm, a, tree);
        if (threadIdx.x == 0) st_xyzz(partial + t0, a);
        __syncthreads();
    }
}

template <class F> __device__ __forceinline__ XYZZ<F> ld_xyzz_cg(const XYZZ<F> *p) {      // written by another CTA of this launch: L2, not L1
    XYZZ<F> v;
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint32_t *d = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 16); i++) { const uint4 x = __ldcg(q + i); d[4 * i] = x.x; d[4 * i + 1] = x.y; d[4 * i + 2] = x.z; d[4 * i + 3] = x.w; }
    return v;
}

// Per bucket region (blockIdx.y; the last one = the "ones" buckets, weight 1) and CTA: chain c owns the buckets [c*seg, (c+1)*seg) and
// computes  S = sum B_j,  T = sum (j - lo + 1) B_j  by running sums, then  T + lo*S; tree over the chains; the last CTA of the region to
// finish adds up the region's CTA sums -> final[region].  counters[region] must be 0 at launch and is 0 again afterwards.
template <class F>
static __global__ void __launch_bounds__(128) msm_reduce_team_kernel(const XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ offsets,
                                                                     uint32_t threads, MsmShape sh, uint32_t seg_ones, uint32_t seg_weighted,
                                                                     XYZZ<F> *__restrict__ cta_out, XYZZ<F> *__restrict__ final_out,
                                                                     uint32_t *__restrict__ counters, int dbg = 0) {
    extern __shared__ uint4 team_sm[];
    __shared__ uint32_t is_last;
    Team<F> tm(team_sm);
    tm.serial = dbg & 4;
    XYZZ<F> *tree = reinterpret_cast<XYZZ<F> *>(reinterpret_cast<char *>(team_sm) + Team<F>::BYTES);
    const uint32_t region = blockIdx.y;
    const bool ones = (region == sh.regions);
    const uint32_t count = ones ? sh.ones : sh.nb;
    const uint32_t base = region * sh.nb;                         // the ones region starts at regions*nb as well
    const uint32_t seg = ones ? seg_ones : seg_weighted;
    const uint32_t lo = (blockIdx.x * TEAM_CHAINS + tm.l) * seg;
    const uint32_t L = msm_range_len(__ldg(offsets + sh.total), threads);
    XYZZ<F> S = XYZZ<F>::inf(), T = XYZZ<F>::inf();
    if (blockIdx.x * TEAM_CHAINS * seg < count) {                 // (CTA-uniform)
        for (uint32_t j = seg; j-- > 0;) {
            const XYZZ<F> o = lo + j < count ? msm_bucket(partial, offsets, L, base + lo + j) : XYZZ<F>::inf();
            team_add(tm, S, o);
            if (!ones && seg > 1) team_add(tm, T, S);
        }
        if (ones) T = S;
        else {
            // weight of bucket j is j + 1:  T (running sums: weights 1..seg inside the segment) + lo * S;  one bucket per chain: (lo + 1) * S
            const uint32_t k = seg > 1 ? lo : lo + 1, kmax = (blockIdx.x * TEAM_CHAINS + TEAM_CHAINS - 1) * seg + (seg > 1 ? 0 : 1);
            if (kmax) {
                const uint32_t kk = lo < count ? k : 0u;
                const XYZZ<F> r = (dbg & 2) ? (kk ? S.mul_small(kk) : XYZZ<F>::inf()) : team_mul_small(tm, S, kk, 31 - __clz(kmax));
                team_add(tm, T, r);
            }
        }
    }
    team_tree_sum(tm, T, tree);
    if (dbg & 1) { if (threadIdx.x == 0) st_xyzz(cta_out + (size_t)region * gridDim.x + blockIdx.x, T); return; }      // host sums the CTA outputs
    if (threadIdx.x == 0) {
        st_xyzz(cta_out + (size_t)region * gridDim.x + blockIdx.x, T);
        __threadfence();
        is_last = atomicAdd(counters + region, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    XYZZ<F> a = XYZZ<F>::inf();
    for (uint32_t k0 = 0; k0 < gridDim.x; k0 += TEAM_CHAINS) {
        const XYZZ<F> o = k0 + tm.l < gridDim.x ? ld_xyzz_cg(cta_out + (size_t)region * gridDim.x + k0 + tm.l) : XYZZ<F>::inf();
        team_add(tm, a, o);
    }
    team_tree_sum(tm, a, tree);
    if (threadIdx.x == 0) { st_xyzz(final_out + region, a); counters[region] = 0; }
}

} // namespace zk
