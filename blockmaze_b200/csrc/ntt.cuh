// Radix-2 number-theoretic transforms over BN254 Fr for sm_100a, plus the evaluation-domain glue of libfqfft that
// BlockMaze's prover reaches:
//   basic_radix2_domain  FFT/iFFT/cosetFFT/icosetFFT/divide_by_Z_on_coset   (libfqfft .../domains/basic_radix2_domain.tcc:26-112,
//                                                                           basic_radix2_domain_aux.tcc:45-79,172-180)
//   step_radix2_domain   idem for m = 2^k + 2^r (mint/redeem: 196608)      (.../domains/step_radix2_domain.tcc:21-248)
//
// Design (B200): a size-2^logn transform is 1-3 shared-memory passes.  Each pass brings a 512..2048-element tile (<= 64 KB, SoA
// by limb so butterflies are bank-conflict free) into shared memory, runs up to 10 butterfly stages there -- two stages at a
// time as radix-4 steps in registers -- and writes it back, i.e. one HBM/L2 round trip per <= 10 stages.  The first pass gathers its input in bit-reversed order (so
// the CLRS decimation-in-time schedule of the reference is kept and the output is in natural order) and can multiply
// by g^i on the way in (coset shift); the last pass can multiply by a per-index power table on the way out (g^-i / n).
// Twiddles omega^j, j < n/2, come from a table built once per domain.
#pragma once
#include <cuda_runtime.h>
#include "ff.cuh"

namespace zk {

// element i is multiplied by lo[i & ((1<<lo_bits)-1)] * hi[i >> lo_bits]  (hi may be null: single-level / constant)
struct PowMul {
    const Fr *lo;
    const Fr *hi;
    int lo_bits;
    __device__ __forceinline__ bool on() const { return lo != nullptr; }
    __device__ __forceinline__ Fr at(uint32_t i) const {
        Fr x = lo[i & ((1u << lo_bits) - 1)];
        if (hi) x = x * hi[i >> lo_bits];
        return x;
    }
};

__device__ __forceinline__ Fr ld_fr(const Fr *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1];
    Fr r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ Fr ldg_fr(const Fr *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr *p, const Fr &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

#ifndef ZK_NTT_INLINE_MUL
#define ZK_NTT_INLINE_MUL 1      // butterfly multiplication inlined (one per loop body; measured 2-3 % faster than the call)
#endif
__device__ __forceinline__ Fr ntt_mul(const Fr &a, const Fr &b) {
#if ZK_NTT_INLINE_MUL
    return Fr::mul_impl(a, b);
#else
    return a * b;
#endif
}
#ifndef ZK_NTT_MINBLOCKS
#define ZK_NTT_MINBLOCKS 2       // CTAs of NTT_MAX_THREADS per SM the register allocation must allow (2 -> 128 registers, 3 -> 80, 4 -> 64)
#endif
// Twiddles are stored PER STAGE, compactly: stage L (butterflies of span 2^L, L = 1..logn) owns the 2^(L-1) powers w_L^j of the primitive
// 2^L-th root at offset 2^(L-1) - 1, so a stage walks its own contiguous run instead of striding through one table of n/2 powers (where
// every stage but the last used 1/2, 1/4, ... of each line it fetched: round 1 measured 1.46 GB of DRAM reads for the 0.5 GB of the last
// pass of a 2^24 transform).  n - 1 elements per direction instead of n/2.
__device__ __forceinline__ const Fr *ntt_tw(const Fr *levels, int stage, uint32_t j) { return levels + ((1u << (stage - 1)) - 1u) + j; }
// levels[2^(L-1) - 1 + j] = flat[j << (logn - L)], flat[j] = w^j (j < n/2)
static __global__ void ntt_tw_levels_kernel(const Fr *__restrict__ flat, int logn, Fr *__restrict__ levels) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;          // position in the level table, 0 .. n-2
    if (e >= (1u << logn) - 1u) return;
    const int L = 32 - __clz(e + 1);                                    // stage: 2^(L-1) <= e + 1 < 2^L
    const uint32_t j = e + 1 - (1u << (L - 1));
    levels[e] = flat[(size_t)j << (logn - L)];
}
constexpr int NTT_MAX_THREADS = 256;
constexpr int NTT_TILE_LOG = 11;          // at most 2048 elements * 32 B = 64 KB of shared memory per CTA; 4 elements per thread

// One pass = stages s0+1 .. s0+k of the decimation-in-time schedule over `n = 2^logn` elements.
//   set  = the 2^k elements that differ only in index bits [s0, s0+k)
//   tile = G = 2^logG sets with consecutive low bits, so global accesses are G*32-byte contiguous runs
template <bool FIRST>
static __global__ void __launch_bounds__(NTT_MAX_THREADS, ZK_NTT_MINBLOCKS)
ntt_pass_kernel(const Fr *__restrict__ src, Fr *__restrict__ dst, const Fr *__restrict__ tw,
                int logn, int s0, int k, int logG, PowMul pre, PowMul post, int last, size_t batch_stride,
                const Fr *__restrict__ sub = nullptr, const Fr *__restrict__ sub_scale = nullptr) {
    extern __shared__ uint32_t sm[];
    src += blockIdx.y * batch_stride; dst += blockIdx.y * batch_stride;      // independent transforms of one launch (A, B, C of the QAP map)
    const int N = 1 << (k + logG);
    const uint32_t set0 = blockIdx.x << logG;
    const uint32_t lowmask = (1u << s0) - 1;

    // The sets of a tile: consecutive ones in general, so that the elements with equal t are contiguous in memory.  The first pass reads
    // through the bit reversal, which turns the TOP bits of the set index into the low address bits: there a tile takes the sets
    // blockIdx.x + g * gridDim.x, and its gather becomes runs of G consecutive elements as well (2^24: 2.15 GB -> 0.5 GB of DRAM reads).
    const uint32_t set_stride = FIRST ? gridDim.x : 1u;
    const uint32_t set_base = FIRST ? blockIdx.x : set0;
    const int nthreads = blockDim.x;
    for (int e = threadIdx.x; e < N; e += nthreads) {
        const int g = e & ((1 << logG) - 1), t = e >> logG;
        const uint32_t set = set_base + g * set_stride;
        const uint32_t addr = ((set >> s0) << (s0 + k)) | ((uint32_t)t << s0) | (set & lowmask);
        Fr x;
        if (FIRST) {
            const uint32_t from = __brev(addr) >> (32 - logn);
            x = ld_fr(src + from);
            if (pre.on()) x = x * pre.at(from);
        } else {
            x = ld_fr(src + addr);
        }
        const int slot = (g << k) | t;
#pragma unroll
        for (int w = 0; w < 8; w++) sm[w * N + slot] = x.v[w];
    }

    // Butterfly stages.  Two stages at a time (radix-4 in registers): a thread takes the four elements that differ in index bits q-1 and
    // q, multiplies by three twiddles (four multiplications, as two radix-2 stages would) and writes them back -- half the shared-memory
    // traffic and half the barriers of a stage-by-stage schedule, and two independent multiplications in flight.  An odd k starts
    // with one radix-2 stage.
    int q = 1;
    if (k & 1) {
        __syncthreads();
        const int half = N >> 1;
        for (int u = threadIdx.x; u < half; u += nthreads) {
            const int g = u >> (k - 1), tt = u & ((1 << (k - 1)) - 1);
            const int i0 = (g << k) | (tt << 1), i1 = i0 + 1;
            const uint32_t j = (set0 + g) & lowmask;                  // tlow = 0 at the first stage
            Fr a, b;
#pragma unroll
            for (int w = 0; w < 8; w++) { a.v[w] = sm[w * N + i0]; b.v[w] = sm[w * N + i1]; }
            if (j != 0) b = ntt_mul(b, ldg_fr(ntt_tw(tw, s0 + 1, j)));
            Fr s = a + b, d = a - b;
#pragma unroll
            for (int w = 0; w < 8; w++) { sm[w * N + i0] = s.v[w]; sm[w * N + i1] = d.v[w]; }
        }
        q = 2;
    }
    const int quarter = N >> 2;
    for (; q < k; q += 2) {
        __syncthreads();
        const int hq = 1 << (q - 1);
        for (int u = threadIdx.x; u < quarter; u += nthreads) {
            const int g = u >> (k - 2), tt = u & ((1 << (k - 2)) - 1);
            const int tlow = tt & (hq - 1);
            const int i0 = (g << k) | ((tt >> (q - 1)) << (q + 1)) | tlow;
            const uint32_t low = (set0 + g) & lowmask;
            const uint32_t j = ((uint32_t)tlow << s0) | low, j2 = ((uint32_t)(tlow + hq) << s0) | low;
            Fr x0, x1, x2, x3;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                x0.v[w] = sm[w * N + i0]; x1.v[w] = sm[w * N + i0 + hq]; x2.v[w] = sm[w * N + i0 + 2 * hq]; x3.v[w] = sm[w * N + i0 + 3 * hq];
            }
            if (j != 0) {                                               // stage q: (x0, x1) and (x2, x3), same twiddle
                const Fr w1 = ldg_fr(ntt_tw(tw, s0 + q, j));
                x1 = ntt_mul(x1, w1); x3 = ntt_mul(x3, w1);
            }
            const Fr a0 = x0 + x1, a1 = x0 - x1;
            Fr a2 = x2 + x3, a3 = x2 - x3;
            if (j != 0) a2 = ntt_mul(a2, ldg_fr(ntt_tw(tw, s0 + q + 1, j)));      // stage q+1: (a0, a2) and (a1, a3)
            a3 = ntt_mul(a3, ldg_fr(ntt_tw(tw, s0 + q + 1, j2)));
            x0 = a0 + a2; x2 = a0 - a2; x1 = a1 + a3; x3 = a1 - a3;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                sm[w * N + i0] = x0.v[w]; sm[w * N + i0 + hq] = x1.v[w]; sm[w * N + i0 + 2 * hq] = x2.v[w]; sm[w * N + i0 + 3 * hq] = x3.v[w];
            }
        }
    }
    __syncthreads();

    for (int e = threadIdx.x; e < N; e += nthreads) {
        int t, g;
        if (FIRST) { t = e & ((1 << k) - 1); g = e >> k; } else { g = e & ((1 << logG) - 1); t = e >> logG; }
        const uint32_t set = set_base + g * set_stride;
        const uint32_t addr = ((set >> s0) << (s0 + k)) | ((uint32_t)t << s0) | (set & lowmask);
        const int slot = (g << k) | t;
        Fr x;
#pragma unroll
        for (int w = 0; w < 8; w++) x.v[w] = sm[w * N + slot];
        if (last && post.on()) x = x * post.at(addr);
        if (last && sub) x = x - ldg_fr(sub + addr) * ldg_fr(sub_scale);      // out = x * post - sub * scale (the QAP map's h = (d - c) / Z, see prover.cu)
        st_fr(dst + addr, x);
    }
}

struct NttPass { int s0, k, logG; };

// split logn into ceil(logn/10) nearly equal passes
static inline int ntt_plan_passes(int logn, NttPass out[4]) {
    int np = (logn + 9) / 10; if (np < 1) np = 1;
    int s0 = 0;
    for (int p = 0; p < np; p++) {
        int k = (logn - s0 + (np - p) - 1) / (np - p);
        // tile = 2^(k+logG) elements: as large as 2048 for big transforms (long contiguous runs), but small enough that a
        // circuit-sized transform still spreads over >= 2 CTAs per SM (2^18: 512 CTAs of 512 elements)
        int tile = logn - 9; if (tile > NTT_TILE_LOG) tile = NTT_TILE_LOG; if (tile < k) tile = k;
        int logG = tile - k;
        if (logG < 0) logG = 0;
        if (k + logG > logn) logG = logn - k;
        if (p > 0 && logG > s0) logG = s0;
        out[p] = NttPass{s0, k, logG};
        s0 += k;
    }
    return np;
}

// dst != src.  tw = per-stage twiddle table (ntt_tw) of omega for a forward transform, of omega^-1 for an inverse one (Montgomery form).
// `batch` transforms, `batch_stride` elements apart in both src and dst, share one launch per pass.
static inline void ntt_launch(cudaStream_t st, const Fr *src, Fr *dst, const Fr *tw, int logn, PowMul pre, PowMul post, int batch = 1,
                              size_t batch_stride = 0, const Fr *sub = nullptr, const Fr *sub_scale = nullptr) {
    NttPass ps[4];
    const int np = ntt_plan_passes(logn, ps);
    for (int p = 0; p < np; p++) {
        const int N = 1 << (ps[p].k + ps[p].logG);
        const size_t smem = (size_t)N * 32;
        const dim3 blocks(1u << (logn - ps[p].k - ps[p].logG), (unsigned)batch);
        const int last = (p == np - 1);
        int threads = N / 4; if (threads < 32) threads = 32; if (threads > NTT_MAX_THREADS) threads = NTT_MAX_THREADS;
        if (p == 0)
            ntt_pass_kernel<true><<<blocks, threads, smem, st>>>(src, dst, tw, logn, ps[p].s0, ps[p].k, ps[p].logG, pre, post, last, batch_stride, sub, sub_scale);
        else
            ntt_pass_kernel<false><<<blocks, threads, smem, st>>>(dst, dst, tw, logn, ps[p].s0, ps[p].k, ps[p].logG, pre, post, last, batch_stride, sub, sub_scale);
    }
}
static inline void ntt_init_attrs() {
    cudaFuncSetAttribute(ntt_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
}

// ---------------------------------------------------------------------------------------------------------------
// table builders (run once per domain)
// out[i] = base^i * scale for i < count, computed as base^(i0) by square-and-multiply per thread chunk of 64
static __global__ void pow_table_kernel(Fr *out, Fr base, Fr scale, uint32_t count, uint64_t stride) {
    const uint32_t chunk = 64;
    uint32_t c0 = (blockIdx.x * blockDim.x + threadIdx.x) * chunk;
    if (c0 >= count) return;
    // step = base^stride ; x = step^c0 * scale
    Fr step = base.pow_u64(stride);
    Fr x = step.pow_u64(c0) * scale;
    for (uint32_t i = c0; i < c0 + chunk && i < count; i++) { st_fr(out + i, x); x = x * step; }
}

// ---------------------------------------------------------------------------------------------------------------
// element-wise glue
// step_radix2 FFT front end (step_radix2_domain.tcc:44-63), in place:
//   a' = a .* pre (coset shift, optional);  c[i] = a'[i] + a'[i+big] (i<small) | a'[i];  d[i] = w^i (a'[i] - a'[i+big] | a'[i])
//   e[i] = sum_j d[i + j*small]   ->   a[0..big) = c, a[big..big+small) = e
static __global__ void step_fft_pre_kernel(Fr *a, const Fr *tw_big2 /* w^i, i < big */, uint32_t big, uint32_t small, PowMul pre, size_t batch_stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= small) return;
    a += blockIdx.y * batch_stride;
    const uint32_t compr = big / small;
    Fr e = Fr::zero();
    for (uint32_t j = 0; j < compr; j++) {
        uint32_t idx = i + j * small;
        Fr x = ld_fr(a + idx);
        if (pre.on()) x = x * pre.at(idx);
        Fr c, d;
        if (j == 0) {
            Fr y = ld_fr(a + idx + big);
            if (pre.on()) y = y * pre.at(idx + big);
            c = x + y; d = x - y;
        } else { c = x; d = x; }
        if (idx != 0) d = d * ldg_fr(tw_big2 + idx);
        e = e + d;
        st_fr(a + idx, c);
    }
    st_fr(a + big + i, e);
}

// step_radix2 iFFT back end (step_radix2_domain.tcc:97-143), in place.  On entry a[0..big) = U0 (already * 1/big),
// a[big..) = U1 (already * 1/small).  post (optional) multiplies the final coefficient i (icosetFFT).
static __global__ void step_ifft_post_kernel(Fr *a, const Fr *tw_big2 /* w^i */, const Fr *tw_big2_inv /* w^-i */, uint32_t big, uint32_t small,
                                      Fr over_two, PowMul post, size_t batch_stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= small) return;
    a += blockIdx.y * batch_stride;
    const uint32_t compr = big / small;
    Fr u0 = ld_fr(a + i), u1 = ld_fr(a + big + i);
    for (uint32_t j = 1; j < compr; j++) {
        uint32_t idx = i + j * small;
        Fr x = ld_fr(a + idx);
        u1 = u1 - x * ldg_fr(tw_big2 + idx);
        if (post.on()) st_fr(a + idx, x * post.at(idx));
    }
    if (i != 0) u1 = u1 * ldg_fr(tw_big2_inv + i);
    Fr lo = (u0 + u1) * over_two, hi = (u0 - u1) * over_two;
    if (post.on()) { lo = lo * post.at(i); hi = hi * post.at(big + i); }
    st_fr(a + i, lo);
    st_fr(a + big + i, hi);
}

// H[i] = (A[i]*B[i] - C[i]) * Zinv(i)  on the coset (r1cs_to_qap.tcc:274-305 with divide_by_Z_on_coset fused).
//   i <  big : zt[i % compr]      (basic domain: big = m, compr = 1)
//   i >= big : z1
// basic domain, C-free variant (see qap_pipeline): only the product of the coset evaluations is needed
static __global__ void qap_product_kernel(Fr *A, const Fr *B, uint32_t m) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) st_fr(A + i, ld_fr(A + i) * ld_fr(B + i));
}
static __global__ void qap_pointwise_kernel(Fr *A, const Fr *B, const Fr *C, uint32_t m, uint32_t big, uint32_t compr, const Fr *zt, Fr z1) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Fr h = ld_fr(A + i) * ld_fr(B + i) - ld_fr(C + i);
    Fr z = (i < big) ? ldg_fr(zt + (i % compr)) : z1;
    st_fr(A + i, h * z);
}
static __global__ void divide_by_z_kernel(Fr *P, uint32_t m, uint32_t big, uint32_t compr, const Fr *zt, Fr z1) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Fr z = (i < big) ? ldg_fr(zt + (i % compr)) : z1;
    st_fr(P + i, ld_fr(P + i) * z);
}

static __global__ void to_mont_kernel(Fr *a, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(a + i, ld_fr(a + i).to_mont());
}
static __global__ void from_mont_kernel(const Fr *a, Fr *out, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_fr(out + i, ld_fr(a + i).from_mont());
}

} // namespace zk
