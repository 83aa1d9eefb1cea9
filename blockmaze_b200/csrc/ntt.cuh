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
#include <cstdlib>
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

constexpr int NTT_DENSE_LOG = 18;           // transforms of at least 2^18 points run with six warps per SM sub-partition (see ntt_pass_kernel)
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
//
// What bounds the kernel (ncu, round 2): with 128 registers four warps share an SM sub-partition, each a chain of dependent instructions,
// so the sub-partition issues ~0.35 instructions per clock whatever they are -- the multiply pipe is only as busy as the share of multiply
// instructions in the stream.  Hence every instruction that is not part of a multiplication counts:
//   * the tile lives in shared memory as two 16-byte halves per element ([half][slot]): 2 loads/stores per element instead of 8;
//   * butterflies are LAZY: values stay in [0, 4r) (4r < 2^256), the Montgomery product skips its final subtraction (a < 4r, twiddle < r
//     gives a result below 2r), u + t is a plain 256-bit addition and u - t + 2r a plain subtraction, and only the u leg is brought below
//     2r first.  One conditional subtraction per butterfly instead of three; intermediate passes store unreduced values, the last pass
//     reduces.  Results are the same field elements, bit for bit after the final reduction.
__device__ __forceinline__ Fr ntt_ld_tile(const uint4 *sm4, int N, int slot) {
    const uint4 a = sm4[slot], b = sm4[N + slot];
    Fr r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void ntt_st_tile(uint4 *sm4, int N, int slot, const Fr &r) {
    sm4[slot] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    sm4[N + slot] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
// (u, v) -> (u + w v, u - w v) on values below 4r; `unit`: the twiddle is 1
__device__ __forceinline__ void ntt_bfly(Fr &u, Fr &v, const Fr &w, bool unit) {
    const Fr t = unit ? Fr::condsub_2p(v) : Fr::mul_lazy(v, w);
    const Fr a = Fr::condsub_2p(u);
    u = Fr::add_lazy(a, t);
    v = Fr::sub_lazy(a, t);
}
// MINB = CTAs of NTT_MAX_THREADS per SM the register allocation must allow: 2 -> up to 128 registers (the kernel needs 92-98: four warps per
// sub-partition), 3 -> 80 registers (six warps; 32 bytes spilled in the first-pass variant).  Measured: 3 is faster from 2^18 up (2^24:
// 3.97 -> 3.59 ms; deposit's QAP map 1.19 -> 1.12 ms), 2 for the smaller transforms of the step domains; 4 (64 registers) loses everywhere.
template <bool FIRST, int MINB>
static __global__ void __launch_bounds__(NTT_MAX_THREADS, MINB)
ntt_pass_kernel(const Fr *__restrict__ src, Fr *__restrict__ dst, const Fr *__restrict__ tw,
                int logn, int s0, int k, int logG, PowMul pre, PowMul post, int last, size_t batch_stride,
                const Fr *__restrict__ sub = nullptr, const Fr *__restrict__ sub_scale = nullptr) {
    extern __shared__ uint4 sm4[];
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
    // A CTA has N/4 threads (at least 32, at most 256): four elements per thread and trip, all four loads issued before the first use so
    // that they share one global-memory latency.
    for (int e0 = threadIdx.x; e0 < N; e0 += 4 * nthreads) {       // (one trip at the circuit sizes; two for the 2048-element tiles)
        Fr x[4]; uint32_t from[4];
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int e = e0 + it * nthreads;
            const int g = e & ((1 << logG) - 1), t = e >> logG;
            const uint32_t set = set_base + g * set_stride;
            const uint32_t addr = ((set >> s0) << (s0 + k)) | ((uint32_t)t << s0) | (set & lowmask);
            from[it] = FIRST ? __brev(addr) >> (32 - logn) : addr;
            if (e < N) x[it] = ld_fr(src + from[it]);
        }
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int e = e0 + it * nthreads;
            if (e < N) {
                if (FIRST && pre.on()) x[it] = Fr::mul_lazy(x[it], pre.at(from[it]));
                const int g = e & ((1 << logG) - 1), t = e >> logG;
                ntt_st_tile(sm4, N, (g << k) | t, x[it]);
            }
        }
    }

    // Butterfly stages.  Two stages at a time (radix-4 in registers): a thread takes the four elements that differ in index bits q-1 and
    // q, multiplies by three twiddles (four multiplications, as two radix-2 stages would) and writes them back -- half the shared-memory
    // traffic and half the barriers of a stage-by-stage schedule.  An odd k starts with one radix-2 stage.
    int q = 1;
    if (k & 1) {
        __syncthreads();
        const int half = N >> 1;
        for (int u = threadIdx.x; u < half; u += nthreads) {
            const int g = u >> (k - 1), tt = u & ((1 << (k - 1)) - 1);
            const int i0 = (g << k) | (tt << 1), i1 = i0 + 1;
            const uint32_t j = (set0 + g) & lowmask;                  // tlow = 0 at the first stage
            Fr a = ntt_ld_tile(sm4, N, i0), b = ntt_ld_tile(sm4, N, i1);
            Fr w = Fr::zero();
            if (j != 0) w = ldg_fr(ntt_tw(tw, s0 + 1, j));
            ntt_bfly(a, b, w, j == 0);
            ntt_st_tile(sm4, N, i0, a); ntt_st_tile(sm4, N, i1, b);
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
            Fr x0 = ntt_ld_tile(sm4, N, i0), x1 = ntt_ld_tile(sm4, N, i0 + hq), x2 = ntt_ld_tile(sm4, N, i0 + 2 * hq), x3 = ntt_ld_tile(sm4, N, i0 + 3 * hq);
            const bool unit = j == 0;
            Fr w = Fr::zero();
            if (!unit) w = ldg_fr(ntt_tw(tw, s0 + q, j));               // stage q: (x0, x1) and (x2, x3), same twiddle
            ntt_bfly(x0, x1, w, unit);
            ntt_bfly(x2, x3, w, unit);
            if (!unit) w = ldg_fr(ntt_tw(tw, s0 + q + 1, j));           // stage q+1: (x0, x2) and (x1, x3)
            ntt_bfly(x0, x2, w, unit);
            ntt_bfly(x1, x3, ldg_fr(ntt_tw(tw, s0 + q + 1, j2)), false);
            ntt_st_tile(sm4, N, i0, x0); ntt_st_tile(sm4, N, i0 + hq, x1); ntt_st_tile(sm4, N, i0 + 2 * hq, x2); ntt_st_tile(sm4, N, i0 + 3 * hq, x3);
        }
    }
    __syncthreads();

    for (int e0 = threadIdx.x; e0 < N; e0 += 4 * nthreads) {
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const int e = e0 + it * nthreads;
            if (e < N) {
                int t, g;
                if (FIRST) { t = e & ((1 << k) - 1); g = e >> k; } else { g = e & ((1 << logG) - 1); t = e >> logG; }
                const uint32_t set = set_base + g * set_stride;
                const uint32_t addr = ((set >> s0) << (s0 + k)) | ((uint32_t)t << s0) | (set & lowmask);
                Fr x = ntt_ld_tile(sm4, N, (g << k) | t);
                if (last) {                                             // the transform's result: fully reduced
                    if (post.on()) { x = Fr::mul_lazy(x, post.at(addr)); x.reduce_once(); }
                    else x = Fr::reduce_4p(x);
                    if (sub) x = x - ldg_fr(sub + addr) * ldg_fr(sub_scale);      // out = x * post - sub * scale (the QAP map's h = (d - c) / Z, see prover.cu)
                }
                st_fr(dst + addr, x);
            }
        }
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
        const bool dense = logn >= NTT_DENSE_LOG;
        const Fr *in = p == 0 ? src : dst;
        if (p == 0 && dense) ntt_pass_kernel<true, 3><<<blocks, threads, smem, st>>>(in, dst, tw, logn, ps[p].s0, ps[p].k, ps[p].logG, pre, post, last, batch_stride, sub, sub_scale);
        else if (p == 0) ntt_pass_kernel<true, 2><<<blocks, threads, smem, st>>>(in, dst, tw, logn, ps[p].s0, ps[p].k, ps[p].logG, pre, post, last, batch_stride, sub, sub_scale);
        else if (dense) ntt_pass_kernel<false, 3><<<blocks, threads, smem, st>>>(in, dst, tw, logn, ps[p].s0, ps[p].k, ps[p].logG, pre, post, last, batch_stride, sub, sub_scale);
        else ntt_pass_kernel<false, 2><<<blocks, threads, smem, st>>>(in, dst, tw, logn, ps[p].s0, ps[p].k, ps[p].logG, pre, post, last, batch_stride, sub, sub_scale);
    }
}
static inline void ntt_init_attrs() {
    cudaFuncSetAttribute(ntt_pass_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt_pass_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt_pass_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(ntt_pass_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
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
