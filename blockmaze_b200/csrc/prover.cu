// GPU pipeline of the Groth16 prover for BlockMaze's circuits (replaces r1cs_gg_ppzksnark_prover,
// libsnark/zk_proof_systems/ppzksnark/r1cs_gg_ppzksnark/r1cs_gg_ppzksnark.tcc:390-506, and everything below it:
// r1cs_to_qap_witness_map r1cs_to_qap.tcc:205-334, libfqfft domains, libff multi_exp).  One DevicePk per (circuit, GPU):
// fixed-base tables, constraint matrices and twiddles stay resident in HBM, and so do a few "lanes" of work buffers so that several
// proofs are in flight; a proof is one H2D copy of the (compact) assignment, ~42 kernel launches on five streams, and a D2H copy of
// a hundred partial sums.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <mutex>
#include <shared_mutex>
#include <thread>
#include "prover.cuh"
#include "pk_format.hpp"
#include "ntt.cuh"
#include "msm_team.cuh"
#include "witness_sha.hpp"

namespace zkp {
using namespace zk;
using zkh::HFr; using zkh::HFq; using zkh::HFq2; using zkh::HG1; using zkh::HG2; using zkh::HG1Affine; using zkh::HG2Affine;

static thread_local int g_launches = 0;     // kernels launched by this thread since the last prove_submit() began
static int g_last_launches = 0;
static thread_local unsigned long long g_h2d_bytes = 0;      // host-to-device bytes of the proof this thread is submitting
static thread_local unsigned long long g_last_bytes[2] = {0, 0};   // host-to-device / device-to-host bytes of the last proof this thread collected
static bool g_isolate_h = false;            // measurement mode, see set_isolate_h()
static const int g_spmv_bits = [] { const char *e = getenv("ZKB200_SPMV_BITS"); return e ? atoi(e) : 1; }();                        // 0: multiply even by 0 and 1
static std::atomic<int> g_device_pending[64];          // proofs submitted and not yet collected, per device, over all resident keys
static const bool g_use_graph = [] { const char *e = getenv("ZKB200_GRAPH"); return !(e && atoi(e) == 0); }();          // 0: enqueue every proof kernel by kernel
static const int g_sort_v2 = [] { const char *e = getenv("ZKB200_SORT_V2"); return e ? atoi(e) : 1; }();
static const int g_full_pow = [] { const char *e = getenv("ZKB200_FULL_POW"); return e ? atoi(e) : 1; }();                           // 0: two-level coset tables in the QAP map
static const bool g_qap_skip_c = [] { const char *e = getenv("ZKB200_QAP_SEVEN"); return !(e && atoi(e) != 0); }();   // ZKB200_QAP_SEVEN=1: the reference's seven transforms
// a launch-configuration failure (shared memory over the limit, bad grid) is not sticky: check it at the launch or the kernel silently does not run
#define ZK_LAUNCH(kernel, grid, block, smem, stream, ...) do { kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); ZK_CUDA(cudaGetLastError()); g_launches++; } while (0)

void cuda_check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) {
        fprintf(stderr, "zkb200: CUDA failure %s: %s -- the B200 prover has no CPU fallback, aborting\n", what, cudaGetErrorString(e));
        abort();
    }
}
int launches_last_prove() { return g_last_launches; }
void transfer_bytes_last_prove(unsigned long long out[2]) { out[0] = g_last_bytes[0]; out[1] = g_last_bytes[1]; }
void set_isolate_h(bool on) { g_isolate_h = on; }
// A proof's kernels, copies and stream joins are captured ONCE per lane into a CUDA graph and replayed with one call (prove_submit).  Events
// the HOST looks at afterwards (completion polls, CUDA-event timings) must then be external record nodes; events that only order streams
// inside the proof stay ordinary captured dependencies.
static thread_local bool g_capturing = false;
static void record_for_host(cudaEvent_t ev, cudaStream_t st) {
    if (g_capturing) ZK_CUDA(cudaEventRecordWithFlags(ev, st, cudaEventRecordExternal));
    else ZK_CUDA(cudaEventRecord(ev, st));
}
// Completion signals the host polls in plain memory: the last thing on a stream bumps a device counter and stores it to a word of pinned
// host memory.  No driver call per poll (cudaEventQuery from two dozen threads is lock traffic inside the driver), and it works the same
// whether the stream's work was enqueued directly or replayed from a graph (the value is not a kernel argument).
static __global__ void signal_kernel(volatile uint32_t *host_flag, uint32_t *counter) {
    const uint32_t v = *counter + 1;
    *counter = v;
    *host_flag = v;
    __threadfence_system();
}
static float elapsed_ms(cudaEvent_t a, cudaEvent_t b) {          // statistics only: never fatal
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { (void)cudaGetLastError(); ms = 0; }
    return ms;
}

static inline Fr to_dev(const HFr &x) { Fr r; memcpy(r.v, x.v, 32); return r; }
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// cudaDeviceSynchronize is illegal while ANY stream of the device is being captured (whatever the capture mode), and keys are loaded lazily
// by caller threads while other threads already prove: a capture holds this lock shared, whoever synchronises the device -- or loads / frees
// a key, with its allocations and default-stream kernels -- holds it exclusively.  Captures take a millisecond, once per lane.
static std::shared_mutex g_capture_mu[64];
static thread_local int t_exclusive = 0;
struct DeviceExclusive {
    int d = 0; bool own;
    DeviceExclusive() { cudaGetDevice(&d); own = (t_exclusive++ == 0); if (own) g_capture_mu[d & 63].lock(); }
    ~DeviceExclusive() { if (own) g_capture_mu[d & 63].unlock(); --t_exclusive; }
};
void device_sync() { DeviceExclusive x; ZK_CUDA(cudaDeviceSynchronize()); }

static bool g_attrs_done[64];
static std::mutex g_attrs_mu;
void device_init(int device) {
    ZK_CUDA(cudaSetDevice(device));
    if (device < 64 && !g_attrs_done[device]) {
        std::lock_guard<std::mutex> lk(g_attrs_mu);
        if (!g_attrs_done[device]) { ntt_init_attrs(); g_attrs_done[device] = true; }
    }
}
std::vector<int> devices_in_use() {
    std::vector<int> v;
    for (int d = 0; d < 64; d++) if (g_attrs_done[d]) v.push_back(d);
    return v;
}

// =====================================================================================================================
// evaluation domains
static HFr root_of_unity(int logn) {     // libff get_root_of_unity (field_utils.tcc:36-51): square the 2^28-th root down
    HFr w; const char *s = "19103219067921713944291392827692070036145651957329286315305642004821462161904";
    HFr::from_dec(s, strlen(s), w);
    for (int i = 28; i > logn; i--) w = w.sqr();
    return w;
}
static int ilog2_ceil(uint64_t n) { int r = (n & (n - 1)) == 0 ? 0 : 1; while (n > 1) { n >>= 1; r++; } return r; }

static void *dev_pow_table(HFr base, HFr scale, uint32_t count, uint64_t stride) {
    Fr *out; ZK_CUDA(cudaMalloc(&out, (size_t)(count ? count : 1) * 32));
    if (count) { pow_table_kernel<<<cdiv(cdiv(count, 64), 128), 128>>>(out, to_dev(base), to_dev(scale), count, stride); ZK_CUDA(cudaGetLastError()); }
    return out;
}
// per-stage twiddle table of a size-2^logn transform (ntt.cuh ntt_tw): built from the flat power table, which is then dropped
static void *dev_tw_levels(HFr w, int logn) {
    const uint32_t n = 1u << logn;
    if (logn == 0) { void *p; ZK_CUDA(cudaMalloc(&p, 32)); return p; }
    void *flat = dev_pow_table(w, HFr::one(), n / 2, 1);
    Fr *levels; ZK_CUDA(cudaMalloc(&levels, (size_t)n * 32));
    ntt_tw_levels_kernel<<<cdiv(n, 256), 256>>>((const Fr *)flat, logn, levels);
    ZK_CUDA(cudaGetLastError());
    device_sync();
    ZK_CUDA(cudaFree(flat));
    return levels;
}
static void *dev_const(const HFr *v, size_t n) {
    void *p; ZK_CUDA(cudaMalloc(&p, n * 32)); ZK_CUDA(cudaMemcpy(p, v, n * 32, cudaMemcpyHostToDevice)); return p;
}

Domain *Domain::build(uint64_t min_size) {
    if (min_size <= 1) return nullptr;
    // get_evaluation_domain.tcc:33-52: basic(min) | step(min) | basic(big+rounded_small) | step(big+rounded_small)
    const int L = ilog2_ceil(min_size);
    uint64_t m = 0; bool step = false;
    if (min_size == (1ull << L)) { m = min_size; }
    else {
        const uint64_t big = 1ull << (L - 1), small = min_size - big, rs = 1ull << ilog2_ceil(small);
        if (small == rs) { m = min_size; step = true; }
        else if (rs == big) { m = 2 * big; }
        else { m = big + rs; step = true; }
    }
    if (ilog2_ceil(m) > 28) return nullptr;
    Domain *d = new Domain();
    d->m = (uint32_t)m; d->step = step;
    const HFr g = HFr::from_u64(5), ginv = g.inverse(), one = HFr::one();
    if (!step) {
        d->big = d->m; d->small = 0; d->log_big = ilog2_ceil(m); d->compr = 1;
    } else {
        d->big = 1u << (ilog2_ceil(m) - 1); d->small = d->m - d->big;
        d->log_big = ilog2_ceil(d->big); d->log_small = ilog2_ceil(d->small); d->compr = d->big / d->small;
    }
    const HFr wb = root_of_unity(d->log_big);
    d->tw_big_f = dev_tw_levels(wb, d->log_big);
    d->tw_big_i = dev_tw_levels(wb.inverse(), d->log_big);
    const HFr big_inv = HFr::from_u64(d->big).inverse(), m_inv = HFr::from_u64(d->m).inverse();
    d->c_big_inv = dev_const(&big_inv, 1);
    d->c_m_inv = dev_const(&m_inv, 1);
    d->over_two = HFr::from_u64(2).inverse();
    const uint32_t nhi = (d->m >> 10) + 1;
    d->g_lo = dev_pow_table(g, one, 1024, 1);
    d->g_hi = dev_pow_table(g, one, nhi, 1024);
    d->g_hi_ninv = dev_pow_table(g, m_inv, nhi, 1024);
    d->gi_lo = dev_pow_table(ginv, one, 1024, 1);
    d->gi_hi = dev_pow_table(ginv, one, nhi, 1024);
    d->gi_hi_ninv = dev_pow_table(ginv, m_inv, nhi, 1024);
    if (!step) {
        // basic_radix2_domain::divide_by_Z_on_coset (basic_radix2_domain.tcc:102-110): Z(g) = g^m - 1
        HFr z = (g.pow64(d->m) - one).inverse();
        d->zt = dev_const(&z, 1); d->z1 = z;
        const HFr z_over_m = z * m_inv;
        d->gi_hi_zninv = dev_pow_table(ginv, z_over_m, nhi, 1024);      // g^-i * Z^-1 / m, high half of the two-level table
        d->c_z_over_m = dev_const(&z_over_m, 1);
        // The QAP map multiplies 3 m elements per proof by such powers; a two-level table costs a second multiplication per element on the
        // pipe that bounds the prover, the flat table one more 32-byte load from L2 (2 x 8 MB for the 2^18 domain: HBM is not what is scarce).
        if (g_full_pow) { d->g_full_ninv = dev_pow_table(g, m_inv, d->m, 1); d->gi_full_zninv = dev_pow_table(ginv, z_over_m, d->m, 1); }
    } else {
        const HFr ws = root_of_unity(d->log_small), om = root_of_unity(d->log_big + 1);
        d->tw_small_f = dev_tw_levels(ws, d->log_small);
        d->tw_small_i = dev_tw_levels(ws.inverse(), d->log_small);
        d->tw_step_f = dev_pow_table(om, one, d->big, 1);
        d->tw_step_i = dev_pow_table(om.inverse(), one, d->big, 1);
        const HFr small_inv = HFr::from_u64(d->small).inverse();
        d->c_small_inv = dev_const(&small_inv, 1);
        // step_radix2_domain::divide_by_Z_on_coset (step_radix2_domain.tcc:221-248)
        const HFr Z0 = g.pow64(d->big) - one;
        const HFr c_sm_Z0 = g.pow64(d->small) * Z0, o_sm_Z0 = om.pow64(d->small) * Z0, o_2sm = om.pow64(2ull * d->small);
        std::vector<HFr> zt(d->compr);
        HFr elt = one;
        for (uint32_t e = 0; e < d->compr; e++) { zt[e] = (c_sm_Z0 * elt - o_sm_Z0).inverse(); elt = elt * o_2sm; }
        d->zt = dev_const(zt.data(), zt.size());
        const HFr go = g * om;
        d->z1 = ((go.pow64(d->big) - one) * (go.pow64(d->small) - om.pow64(d->small))).inverse();
    }
    device_sync();
    return d;
}
void Domain::release() {
    void *ps[] = {tw_big_f, tw_big_i, tw_small_f, tw_small_i, tw_step_f, tw_step_i, g_lo, g_hi, g_hi_ninv, gi_lo, gi_hi, gi_hi_ninv,
                  c_big_inv, c_small_inv, c_m_inv, zt, gi_hi_zninv, c_z_over_m, g_full_ninv, gi_full_zninv};
    for (void *p : ps) if (p) cudaFree(p);
}

static PowMul pm_none() { return PowMul{nullptr, nullptr, 0}; }
static PowMul pm_const(const void *c) { return PowMul{(const Fr *)c, nullptr, 0}; }
static PowMul pm_two(const void *lo, const void *hi) { return PowMul{(const Fr *)lo, (const Fr *)hi, 10}; }
static PowMul pm_full(const void *t) { return PowMul{(const Fr *)t, nullptr, 31}; }             // one entry per index

static void ntt(cudaStream_t st, const void *src, void *dst, const void *tw, int logn, PowMul pre, PowMul post, int batch = 1, size_t stride = 0,
                const void *sub = nullptr, const void *sub_scale = nullptr) {
    if (logn == 0) {
        for (int b = 0; b < batch; b++) ZK_CUDA(cudaMemcpyAsync((Fr *)dst + b * stride, (const Fr *)src + b * stride, 32, cudaMemcpyDeviceToDevice, st));
        return;
    }
    NttPass ps[4]; g_launches += ntt_plan_passes(logn, ps);
    ntt_launch(st, (const Fr *)src, (Fr *)dst, (const Fr *)tw, logn, pre, post, batch, stride, (const Fr *)sub, (const Fr *)sub_scale);
    ZK_CUDA(cudaGetLastError());
}

// inverse transform src -> dst (dst != src), coefficient i additionally multiplied by `post` (e.g. g^i for the coset shift
// that follows, or g^-i for icosetFFT).  For the basic domain the 1/m factor must be folded into `post` by the caller.
// `batch` independent vectors, `stride` elements apart in src and in dst, go through every launch together.
static void domain_ifft(cudaStream_t st, const Domain &d, void *src, void *dst, PowMul post_basic, PowMul post_step, int batch = 1, size_t stride = 0) {
    Fr *s = (Fr *)src, *t = (Fr *)dst;
    if (!d.step) { ntt(st, s, t, d.tw_big_i, d.log_big, pm_none(), post_basic, batch, stride); return; }
    ntt(st, s, t, d.tw_big_i, d.log_big, pm_none(), pm_const(d.c_big_inv), batch, stride);
    ntt(st, s + d.big, t + d.big, d.tw_small_i, d.log_small, pm_none(), pm_const(d.c_small_inv), batch, stride);
    ZK_LAUNCH(step_ifft_post_kernel, dim3(cdiv(d.small, 128), batch), 128, 0, st, t, (const Fr *)d.tw_step_f, (const Fr *)d.tw_step_i, d.big, d.small,
              to_dev(d.over_two), post_step, stride);
}
// forward transform src -> dst (dst != src; src is clobbered for the step domain), input coefficient i first multiplied by `pre`
static void domain_fft(cudaStream_t st, const Domain &d, void *src, void *dst, PowMul pre, int batch = 1, size_t stride = 0) {
    Fr *s = (Fr *)src, *t = (Fr *)dst;
    if (!d.step) { ntt(st, s, t, d.tw_big_f, d.log_big, pre, pm_none(), batch, stride); return; }
    ZK_LAUNCH(step_fft_pre_kernel, dim3(cdiv(d.small, 128), batch), 128, 0, st, s, (const Fr *)d.tw_step_f, d.big, d.small, pre, stride);
    ntt(st, s, t, d.tw_big_f, d.log_big, pm_none(), pm_none(), batch, stride);
    ntt(st, s + d.big, t + d.big, d.tw_small_f, d.log_small, pm_none(), pm_none(), batch, stride);
}

void domain_op(cudaStream_t st, const Domain &d, int op, void *data, void *tmp) {
    const size_t bytes = (size_t)d.m * 32;
    switch (op) {
    case OP_FFT:
    case OP_COSET_FFT:
        ZK_CUDA(cudaMemcpyAsync(tmp, data, bytes, cudaMemcpyDeviceToDevice, st));
        domain_fft(st, d, tmp, data, op == OP_COSET_FFT ? pm_two(d.g_lo, d.g_hi) : pm_none());
        break;
    case OP_IFFT:
        ZK_CUDA(cudaMemcpyAsync(tmp, data, bytes, cudaMemcpyDeviceToDevice, st));
        domain_ifft(st, d, tmp, data, pm_const(d.c_m_inv), pm_none());
        break;
    case OP_ICOSET_FFT:
        ZK_CUDA(cudaMemcpyAsync(tmp, data, bytes, cudaMemcpyDeviceToDevice, st));
        domain_ifft(st, d, tmp, data, pm_two(d.gi_lo, d.gi_hi_ninv), pm_two(d.gi_lo, d.gi_hi));
        break;
    case OP_DIVIDE_BY_Z:
        ZK_LAUNCH(divide_by_z_kernel, cdiv(d.m, 256), 256, 0, st, (Fr *)data, d.m, d.big, d.compr, (const Fr *)d.zt, to_dev(d.z1));
        break;
    }
}

// =====================================================================================================================
// sparse A.w / B.w / C.w  (linear_combination::evaluate, libsnark/relations/variable.tcc:262; r1cs_to_qap.tcc:227-236,281-285)
// SPMV_G lanes per constraint row (97 % of the rows have <= 4 terms, but the packing constraints of the SHA-256 gadgets have
// 33-65 and would serialise a whole warp behind one lane); partial sums meet in a shuffle tree.  Coefficient dictionary
// index 0 = +1, 1 = -1 (no multiplication).
#ifndef ZK_SPMV_G
#define ZK_SPMV_G 4
#endif
constexpr int SPMV_G = ZK_SPMV_G;
struct SpmvArgs { const uint32_t *rowptr[3], *col[3], *coef[3]; Fr *out[3]; };
// The launch covers all m rows of the evaluation vectors: rows past the constraints are the zero padding of the domain, except that aA
// continues with (1, w_1 .. w_inputs) (r1cs_to_qap.tcc:227-230) -- no separate memsets or copy kernel.
__global__ void __launch_bounds__(128) spmv_kernel(SpmvArgs A, const Fr *__restrict__ dict, const Fr *__restrict__ w, uint32_t rows, uint32_t m,
                                                   uint32_t num_inputs, int bits) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = tid / SPMV_G, sub = tid % SPMV_G;
    const uint32_t *__restrict__ rowptr = A.rowptr[blockIdx.y], *__restrict__ col = A.col[blockIdx.y], *__restrict__ coef = A.coef[blockIdx.y];
    Fr acc = Fr::zero();
    if (i < rows) {
        for (uint32_t k = __ldg(rowptr + i) + sub, e = __ldg(rowptr + i + 1); k < e; k += SPMV_G) {
            const uint32_t ci = __ldg(coef + k);
            const Fr x = ldg_fr(w + __ldg(col + k));
            // 51 % of a BlockMaze assignment is 0 and 45 % is 1 (the bits of the SHA-256 gadgets), and most dictionary coefficients are the
            // powers of two of the packing constraints, which multiply exactly those bits: 0 * c adds nothing, 1 * c is c.  The
            // multiplication is left for the few products of two wide values.
            if (bits && x.is_zero()) continue;
            if (ci == 0) acc = acc + x;
            else if (ci == 1) acc = acc - x;
            else {
                const Fr cf = ldg_fr(dict + ci);
                if (bits && x == Fr::one()) acc = acc + cf;
                else acc = acc + x * cf;
            }
        }
    }
#pragma unroll
    for (int d = SPMV_G / 2; d > 0; d >>= 1) {
        Fr o;
#pragma unroll
        for (int q = 0; q < 8; q++) o.v[q] = __shfl_down_sync(0xffffffffu, acc.v[q], d, SPMV_G);
        acc = acc + o;
    }
    if (i < rows) { if (sub == 0) st_fr(A.out[blockIdx.y] + i, acc); }
    else if (i < m && sub == 0) st_fr(A.out[blockIdx.y] + i, (blockIdx.y == 0 && i - rows <= num_inputs) ? ldg_fr(w + (i - rows)) : Fr::zero());
}
// r1cs_constraint_system::is_satisfied (r1cs.tcc:133-164) on the evaluation vectors: flag |= (A_i * B_i != C_i)
__global__ void sat_check_kernel(const Fr *A, const Fr *B, const Fr *C, uint32_t nc, uint32_t *flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    if (ld_fr(A + i) * ld_fr(B + i) != ld_fr(C + i)) atomicOr(flag, 1u);
}

// =====================================================================================================================
// proving-key point decompression (alt_bn128_g1.cpp:420-465, alt_bn128_g2.cpp:433-478) -- one thread per point
struct SqrtConsts { uint32_t e_sqrt[8]; Fq half; Fq2 twist_b; Fq b; };   // (q+1)/4, 1/2, b' = 3/(9+u), b = 3
__device__ __forceinline__ Fq fq_sqrt(const Fq &a, const uint32_t e[8], bool &ok) { Fq r = a.pow(e); ok = (r.sqr() == a); return r; }

__global__ void decompress_g1_kernel(const zkpk::CompressedG1 *in, uint32_t n, SqrtConsts k, Affine<Fq> *out, uint8_t *skip, uint32_t *bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const zkpk::CompressedG1 c = in[i];
    Affine<Fq> p = Affine<Fq>::inf();
    const bool inf = (c.flags >> 1) & 1;
    if (!inf) {
        Fq x; memcpy(x.v, c.x, 32);
        bool ok; Fq y = fq_sqrt(x.sqr() * x + k.b, k.e_sqrt, ok);
        if (!ok) atomicAdd(bad, 1u);
        if ((y.from_mont().v[0] & 1) != (uint32_t)(c.flags & 1)) y = y.neg();
        p.x = x; p.y = y;
    }
    out[i] = p;
    if (skip) skip[i] = inf ? 1 : 0;
}
__global__ void decompress_g2_kernel(const zkpk::CompressedG2 *in, uint32_t n, SqrtConsts k, Affine<Fq2> *out, uint32_t *bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const zkpk::CompressedG2 c = in[i];
    Affine<Fq2> p = Affine<Fq2>::inf();
    if (!((c.flags >> 1) & 1)) {
        Fq2 x; memcpy(x.c0.v, c.x, 32); memcpy(x.c1.v, c.x + 32, 32);
        const Fq2 a = x.sqr() * x + k.twist_b;
        Fq2 y; bool ok;
        if (a.c1.is_zero()) {
            Fq r0 = fq_sqrt(a.c0, k.e_sqrt, ok);
            if (ok) { y.c0 = r0; y.c1 = Fq::zero(); }
            else { y.c0 = Fq::zero(); y.c1 = fq_sqrt(a.c0.neg(), k.e_sqrt, ok); }
        } else {
            const Fq nrm = fq_sqrt(a.c0.sqr() + a.c1.sqr(), k.e_sqrt, ok);
            Fq x0 = Fq::zero();
            if (ok) {
                x0 = fq_sqrt((a.c0 + nrm) * k.half, k.e_sqrt, ok);
                if (!ok) x0 = fq_sqrt((a.c0 - nrm) * k.half, k.e_sqrt, ok);
            }
            y.c0 = x0; y.c1 = a.c1 * x0.dbl().inverse();
            if (ok) ok = (y.sqr() == a);
        }
        if (!ok) atomicAdd(bad, 1u);
        if ((y.c0.from_mont().v[0] & 1) != (uint32_t)(c.flags & 1)) y = y.neg();
        p.x = x; p.y = y;
    }
    out[i] = p;
}

static SqrtConsts sqrt_consts() {
    SqrtConsts k;
    uint64_t e[4]; memcpy(e, zkh::FqTag::MOD, 32);
    e[0] += 1;                                   // q + 1 (no carry: low limb of q is not all ones)
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 2) | (i < 3 ? e[i + 1] << 62 : 0);
    memcpy(k.e_sqrt, e, 32);
    HFq half = HFq::from_u64(2).inverse(); memcpy(k.half.v, half.v, 32);
    HFq2 nine_u{HFq::from_u64(9), HFq::one()};
    HFq2 tb = HFq2{HFq::from_u64(3), HFq::zero()} * nine_u.inverse();
    memcpy(k.twist_b.c0.v, tb.c0.v, 32); memcpy(k.twist_b.c1.v, tb.c1.v, 32);
    HFq b = HFq::from_u64(3); memcpy(k.b.v, b.v, 32);
    return k;
}

template <class C, class A>
static void decompress_vec(const std::vector<C> &v, A **out, uint8_t **skip, bool g2, uint32_t *d_bad, const SqrtConsts &k) {
    const uint32_t n = (uint32_t)v.size();
    C *d_in; ZK_CUDA(cudaMalloc(&d_in, sizeof(C) * (size_t)(n ? n : 1)));
    ZK_CUDA(cudaMemcpy(d_in, v.data(), sizeof(C) * (size_t)n, cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMalloc(out, sizeof(A) * (size_t)(n + 8)));
    if (skip) ZK_CUDA(cudaMalloc(skip, n + 8));
    if (n) {
        if constexpr (sizeof(A) == 64) decompress_g1_kernel<<<cdiv(n, 128), 128>>>((const zkpk::CompressedG1 *)d_in, n, k, (Affine<Fq> *)*out, skip ? *skip : nullptr, d_bad);
        else decompress_g2_kernel<<<cdiv(n, 128), 128>>>((const zkpk::CompressedG2 *)d_in, n, k, (Affine<Fq2> *)*out, d_bad);
        ZK_CUDA(cudaGetLastError());
    }
    device_sync();
    ZK_CUDA(cudaFree(d_in));
    (void)g2;
}

// =====================================================================================================================
// MSM
void MsmPlan::init(uint32_t n_, int c_, uint32_t ones_, bool g1, bool g2, bool expanded_, int affine_rounds_) {
    n = n_; c = c_; windows = (255 + c - 1) / c; nb = 1u << (c - 1); ones = ones_; expanded = expanded_;
    affine_rounds = (g1 && !g2 && affine_rounds_ > 0) ? (affine_rounds_ > 4 ? 4 : affine_rounds_) : 0;
    if (const char *e = getenv("ZKB200_AFFINE_ALWAYS")) affine_always = atoi(e) != 0;
    regions = expanded ? 1u : (uint32_t)windows;
    total = regions * nb + ones;
    seg = 4; if (seg > nb) seg = nb;
    // The 8-bit windows of the witness queries leave 128 weighted buckets: one per thread there (j * B_j by double-and-add, <= 13 point
    // operations) instead of running sums over 4 plus a 7-bit multiple -- the reduction is a chain of dependent additions, 19 us each in G2.
    seg_weighted = nb <= 256 ? 1 : seg;
    const uint32_t need_w = cdiv(cdiv(nb, seg_weighted), MSM_RED_THREADS), need_o = cdiv(cdiv(ones ? ones : 1, seg), MSM_RED_THREADS);
    bpw = need_w > need_o ? need_w : need_o;
    ZK_CUDA(cudaMalloc(&counts, (size_t)(total + 1) * 4));
    ZK_CUDA(cudaMalloc(&offsets, (size_t)(total + 1) * 4));
    ZK_CUDA(cudaMalloc(&cursors, (size_t)(total + 1) * 4));
    ZK_CUDA(cudaMemset(counts, 0, (size_t)(total + 1) * 4));           // sort v2 leaves the counts cleared after every run
    entries_cap = (size_t)n * windows + 16;
    if (affine_rounds) entries_cap += (size_t)total * ((1u << affine_rounds) - 1);        // every non-empty bucket padded to a multiple of 2^rounds
    ZK_CUDA(cudaMalloc(&entries, entries_cap * 4));
    ZK_CUDA(cudaMalloc(&heavy, (size_t)(MSM_HEAVY_MAX + 1) * 4)); ZK_CUDA(cudaMalloc(&heavy_g2, (size_t)(MSM_HEAVY_MAX + 1) * 4));
    ZK_CUDA(cudaEventCreateWithFlags(&ev_sorted, cudaEventDisableTiming));
    // one resident wave of the accumulate kernel: every thread then gets the same share of the sorted entries
    int dev = 0, sms = 0, occ1 = 0, occ2 = 0;
    ZK_CUDA(cudaGetDevice(&dev)); ZK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ZK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, msm_accumulate_kernel<Fq>, 128, 0));
    ZK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, msm_accumulate_kernel<Fq2>, 128, 0));
    if (const char *e = getenv("ZKB200_ACC_CTAS")) { const int v = atoi(e); if (v > 0 && v < occ1) occ1 = v; if (v > 0 && v < occ2) occ2 = v; }
    // A proof that has the GPU to itself cuts the entry list into `waves_alone` waves of shorter ranges: dynamic CTA placement then evens
    // out SM-level variance and the accumulate kernel is 5-10 % faster, at the price of more pieces to fold.  With other proofs in flight
    // the extra fold work costs more than the tail it removes (those proofs fill the tail anyway), so one wave is used.  G1 only.
    waves_alone = 2; if (const char *e = getenv("ZKB200_ACC_WAVES")) { waves_alone = atoi(e); if (waves_alone < 1) waves_alone = 1; if (waves_alone > 4) waves_alone = 4; }
    acc_threads_g1 = (uint32_t)(sms * (occ1 > 0 ? occ1 : 1) * 128);
    acc_threads_g2 = (uint32_t)(sms * (occ2 > 0 ? occ2 : 1) * 128);
    ZK_CUDA(cudaEventCreate(&ev_acc0)); ZK_CUDA(cudaEventCreate(&ev_acc1));
    if (affine_rounds) {
        int per_sm = 2;                                              // CTAs of 128 threads per SM for the pair rounds: long batches beat occupancy
        if (const char *e = getenv("ZKB200_AFF_CTAS")) { const int v = atoi(e); if (v > 0 && v <= 16) per_sm = v; }
        aff_threads = (uint32_t)(sms * per_sm * 128);
        ZK_CUDA(cudaMalloc(&offsets_shifted, (size_t)(total + 1) * 4));
        size_t items = entries_cap;
        for (int r = 0; r < affine_rounds; r++) { items = items / 2 + 1; ZK_CUDA(cudaMalloc(&aff_pts[r], items * sizeof(G1Affine))); }
        ZK_CUDA(cudaMalloc(&aff_scratch, (entries_cap / 2 + 1) * 32));
    }
    team = 7; if (const char *e = getenv("ZKB200_TEAM")) team = atoi(e) & 7;          // bit 0: fold, bit 1: heavy fold, bit 2: reduction
    if (const char *e = getenv("ZKB200_TEAM_ALWAYS")) team_always = atoi(e) != 0;
    if (const char *e = getenv("ZKB200_TEAM_DBG")) team_dbg = atoi(e);
    seg_team = nb <= 256 ? 1 : 4;
    { const uint32_t cw = cdiv(cdiv(nb, seg_team), TEAM_CHAINS), co = cdiv(cdiv(ones ? ones : 1, seg), TEAM_CHAINS); bpw_team = cw > co ? cw : co; }
    ZK_CUDA(cudaMalloc(&team_counters, (size_t)(regions + 1) * 8));               // G1 half, G2 half: the two run side by side for the B query
    ZK_CUDA(cudaMemset(team_counters, 0, (size_t)(regions + 1) * 8));
    const size_t nout = (size_t)(regions + 1) * (bpw > bpw_team ? bpw : bpw_team);
    if (g1) ZK_CUDA(cudaMalloc(&final_g1, (size_t)(regions + 1) * sizeof(G1XYZZ)));
    if (g2) ZK_CUDA(cudaMalloc(&final_g2, (size_t)(regions + 1) * sizeof(G2XYZZ)));
    if (g1) {
        ZK_CUDA(cudaMalloc(&buckets_g1, ((size_t)acc_threads_g1 * waves_alone + total + 1) * sizeof(G1XYZZ)));
        ZK_CUDA(cudaMalloc(&out_g1, nout * sizeof(G1XYZZ)));
        ZK_CUDA(cudaMallocHost(&h_out_g1, nout * sizeof(G1XYZZ)));
    }
    if (g2) {
        ZK_CUDA(cudaMalloc(&buckets_g2, (size_t)(acc_threads_g2 + total + 1) * sizeof(G2XYZZ)));
        ZK_CUDA(cudaMalloc(&out_g2, nout * sizeof(G2XYZZ)));
        ZK_CUDA(cudaMallocHost(&h_out_g2, nout * sizeof(G2XYZZ)));
    }
}
float MsmPlan::last_acc_ms() const { float ms = 0; if (ev_acc0 && cudaEventElapsedTime(&ms, ev_acc0, ev_acc1) != cudaSuccess) { (void)cudaGetLastError(); ms = 0; } return ms; }
void MsmPlan::release() {
    void *ps[] = {counts, offsets, cursors, entries, buckets_g1, buckets_g2, out_g1, out_g2, heavy, heavy_g2, offsets_shifted, aff_pts[0], aff_pts[1], aff_pts[2],
                  aff_pts[3], aff_scratch, team_counters, final_g1, final_g2};
    if (ev_sorted) cudaEventDestroy(ev_sorted);
    for (void *p : ps) if (p) cudaFree(p);
    if (ev_acc0) cudaEventDestroy(ev_acc0);
    if (ev_acc1) cudaEventDestroy(ev_acc1);
    if (h_out_g1) cudaFreeHost(h_out_g1);
    if (h_out_g2) cudaFreeHost(h_out_g2);
    *this = MsmPlan();
}

// out[k*n + i] = 2^(c*k) * in[i]  (device -> device), the fixed-base table of an `expanded` plan
void *msm_expand_bases(const void *bases, uint32_t n, int c, bool g2) {
    const int windows = (255 + c - 1) / c;
    void *out;
    const size_t pt = g2 ? sizeof(G2Affine) : sizeof(G1Affine);
    ZK_CUDA(cudaMalloc(&out, (size_t)windows * (n ? n : 1) * pt));
    if (n) {
        if (g2) msm_expand_bases_kernel<Fq2><<<cdiv(n, 128), 128>>>((const G2Affine *)bases, n, c, windows, (G2Affine *)out);
        else msm_expand_bases_kernel<Fq><<<cdiv(n, 128), 128>>>((const G1Affine *)bases, n, c, windows, (G1Affine *)out);
        ZK_CUDA(cudaGetLastError());
    }
    return out;
}

template <class F>
static void msm_points(cudaStream_t st, MsmPlan &p, const MsmShape &sh, const Affine<F> *bases, XYZZ<F> *partial, XYZZ<F> *out, void *h_out, bool timed) {
    const uint32_t *off = (const uint32_t *)p.offsets;
    const uint32_t *entries = (const uint32_t *)p.entries;
    if (timed) record_for_host(p.ev_acc0, st);
    if constexpr (sizeof(F) == 32) {
        if (p.rounds_now) {
            // halving rounds: list r holds the pairwise sums of list r-1 (list -1 = the sorted entries pointing into the base table)
            const Affine<Fq> *in = bases;
            for (int r = 0; r < p.rounds_now; r++) {
                Affine<Fq> *o = (Affine<Fq> *)p.aff_pts[r];
                if (r == 0) ZK_LAUNCH(msm_affine_round_kernel<true>, p.aff_threads / 128, 128, 0, st, in, entries, off, p.total, 0, o, (Fq *)p.aff_scratch, p.aff_threads);
                else ZK_LAUNCH(msm_affine_round_kernel<false>, p.aff_threads / 128, 128, 0, st, in, (const uint32_t *)nullptr, off, p.total, r, o, (Fq *)p.aff_scratch, p.aff_threads);
                in = o;
            }
            bases = in; entries = nullptr; off = (const uint32_t *)p.offsets_shifted;
        }
    }
    const uint32_t T = sizeof(F) == 32 ? p.acc_threads_g1 * (p.alone ? (uint32_t)p.waves_alone : 1u) : p.acc_threads_g2;
    uint32_t *heavy = (uint32_t *)(sizeof(F) == 32 ? p.heavy : p.heavy_g2);
    if (!p.sort_v2_now) ZK_CUDA(cudaMemsetAsync(heavy, 0, 4, st));          // (sort v2: cleared by the scan kernel)
    ZK_LAUNCH(msm_accumulate_kernel<F>, T / 128, 128, 0, st, bases, off, entries, p.total, T, partial);
    if (timed) record_for_host(p.ev_acc1, st);
    // Tails.  With the GPU to itself (one proof: latency is what counts) a proof runs them on team point operations (msm_team.cuh: 32 chains per
    // CTA, the four warps share every point addition, 2.0-2.5x shorter chains; the reduction also returns ONE point per region).  With other
    // proofs in flight the one-thread-per-chain kernels are used: the tails then hide under the other proofs' kernels anyway, and the team
    // kernels' 4x threads and barriers cost 5-13 % of pipelined throughput (profiles/r02_notes.md).  The piece fold of a big bucket set (the H
    // query: 32768 buckets, 150-230 k pieces) is bound by throughput, not by latency, and stays on the one-thread kernel as well.
    const size_t tsm = team_smem_bytes<F>();
    const int team = p.team_now;
    if ((team & 1) && p.total <= MSM_TEAM_FOLD_MAX_BUCKETS) ZK_LAUNCH(msm_fold_small_team_kernel<F>, cdiv(p.total, TEAM_CHAINS), 128, tsm, st, partial, off, p.total, T, heavy);
    else ZK_LAUNCH(msm_fold_small_kernel<F>, cdiv(p.total, 128), 128, 0, st, partial, off, p.total, T, heavy);
    if (team & 2) ZK_LAUNCH(msm_fold_heavy_team_kernel<F>, 64, 128, tsm, st, partial, off, p.total, T, (const uint32_t *)heavy);
    else ZK_LAUNCH(msm_fold_heavy_kernel<F>, 64, MSM_HEAVY_THREADS, MSM_HEAVY_THREADS * sizeof(XYZZ<F>), st, partial, off, p.total, T, (const uint32_t *)heavy);
    if (team & 4) {
        XYZZ<F> *fin = (XYZZ<F> *)(sizeof(F) == 32 ? p.final_g1 : p.final_g2);
        ZK_LAUNCH(msm_reduce_team_kernel<F>, dim3(p.bpw_team, sh.regions + 1), 128, tsm, st, (const XYZZ<F> *)partial, off, T, sh, p.seg, p.seg_team, out, fin,
                  (uint32_t *)p.team_counters + (sizeof(F) == 32 ? 0 : sh.regions + 1), p.team_dbg);
        if (p.team_dbg & 1) ZK_CUDA(cudaMemcpyAsync(h_out, out, (size_t)(sh.regions + 1) * p.bpw_team * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, st));
        else ZK_CUDA(cudaMemcpyAsync(h_out, fin, (size_t)(sh.regions + 1) * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, st));
        return;
    }
    const dim3 rgrid(p.bpw, sh.regions + 1);
    ZK_LAUNCH(msm_reduce_kernel<F>, rgrid, MSM_RED_THREADS, MSM_RED_THREADS * sizeof(XYZZ<F>), st, (const XYZZ<F> *)partial, off, T, sh, p.seg, p.seg_weighted, p.bpw, out);
    ZK_CUDA(cudaMemcpyAsync(h_out, out, (size_t)(sh.regions + 1) * p.bpw * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, st));
}

// what the next run of the plan does, from the flags prove_submit set (alone, latency); also called when a captured proof is replayed, so
// that msm_finish reads the partial sums the way the replayed kernels wrote them
static void msm_set_mode(MsmPlan &p) {
    // A proof that has the GPU to itself skips the affine rounds: with nothing else resident they are bound by gather latency, not by the
    // multiply pipe, and the plain XYZZ accumulation finishes sooner; with other proofs in flight their 40 % fewer multiplications win.
    p.rounds_now = (p.affine_rounds && (!p.alone || p.affine_always)) ? p.affine_rounds : 0;
    p.team_now = (p.latency || p.team_always) ? p.team : 0;
    p.sort_v2_now = g_sort_v2 && (p.c == 16 || p.c == 8);
}
void msm_run(cudaStream_t st, MsmPlan &p, ScalarRef sc, const uint8_t *skip, const void *bases_g1, const void *bases_g2, cudaStream_t st_g2) {
    const MsmShape sh = msm_shape(p.n, p.c, p.ones, p.expanded ? 1 : 0);
    ScalarSrc src{(const uint32_t *)sc.scalars, sc.map, sc.offset, sc.montgomery};
    // sort v2 (default; ZKB200_SORT_V2=0 selects the first version): the scan seeds absolute cursors, clears the counts for the next run and
    // the queues of oversized buckets, and the scatter keeps all atomics of a scalar in flight -- no memsets, shorter scatter
    msm_set_mode(p);
    const bool v2 = p.sort_v2_now;
    if (!v2) {
        ZK_CUDA(cudaMemsetAsync(p.counts, 0, (size_t)(p.total + 1) * 4, st));
        ZK_CUDA(cudaMemsetAsync(p.cursors, 0, (size_t)(p.total + 1) * 4, st));
    }
    if (p.n) ZK_LAUNCH(msm_count_kernel, cdiv(p.n, 256), 256, 0, st, src, skip, sh, (uint32_t *)p.counts);
    if (p.rounds_now) ZK_CUDA(cudaMemsetAsync(p.entries, 0xff, p.entries_cap * 4, st));          // pads of the bucket runs = null entries
    ZK_LAUNCH(msm_scan_kernel, 1, 1024, 0, st, (uint32_t *)p.counts, (uint32_t *)p.offsets, p.total, p.rounds_now,
              (uint32_t *)(p.rounds_now ? p.offsets_shifted : nullptr), (uint32_t *)(v2 ? p.cursors : nullptr), (uint32_t *)(v2 ? p.heavy : nullptr),
              (uint32_t *)(v2 ? p.heavy_g2 : nullptr));
    if (p.n && v2 && p.c == 16) ZK_LAUNCH(msm_scatter_abs_kernel<16>, cdiv(p.n, 256), 256, 0, st, src, skip, sh, (uint32_t *)p.cursors, (uint32_t *)p.entries);
    else if (p.n && v2) ZK_LAUNCH(msm_scatter_abs_kernel<8>, cdiv(p.n, 256), 256, 0, st, src, skip, sh, (uint32_t *)p.cursors, (uint32_t *)p.entries);
    else if (p.n) ZK_LAUNCH(msm_scatter_kernel, cdiv(p.n, 256), 256, 0, st, src, skip, sh, (const uint32_t *)p.offsets, (uint32_t *)p.cursors, (uint32_t *)p.entries);
    // the G2 half of a knowledge-commitment query shares the digit sort and then runs beside the G1 half on its own stream
    if (bases_g2 && st_g2) { ZK_CUDA(cudaEventRecord(p.ev_sorted, st)); ZK_CUDA(cudaStreamWaitEvent(st_g2, p.ev_sorted, 0)); }
    if (bases_g2) msm_points<Fq2>(st_g2 ? st_g2 : st, p, sh, (const G2Affine *)bases_g2, (G2XYZZ *)p.buckets_g2, (G2XYZZ *)p.out_g2, p.h_out_g2, false);
    if (bases_g1) msm_points<Fq>(st, p, sh, (const G1Affine *)bases_g1, (G1XYZZ *)p.buckets_g1, (G1XYZZ *)p.out_g1, p.h_out_g1, true);
}

template <class P> static P msm_finish(const MsmPlan &p, const void *h_out) {
    const P *part = (const P *)h_out;
    const uint32_t bpw = (p.team_now & 4) ? ((p.team_dbg & 1) ? p.bpw_team : 1u) : p.bpw;                                   // team tails: the regions were summed on the device
    P acc = P::inf();
    if (p.expanded) {
        for (uint32_t b = 0; b < 2 * bpw; b++) acc = acc.add(part[b]);         // the single bucket region, then the ones region
        return acc;
    }
    for (int w = p.windows - 1; w >= 0; w--) {                                  // windowed layout: Horner over the per-window sums
        for (int i = 0; i < p.c; i++) acc = acc.dbl();
        for (uint32_t b = 0; b < bpw; b++) acc = acc.add(part[(size_t)w * bpw + b]);
    }
    for (uint32_t b = 0; b < bpw; b++) acc = acc.add(part[(size_t)p.windows * bpw + b]);
    return acc;
}
HG1 msm_finish_g1(const MsmPlan &p) { return msm_finish<HG1>(p, p.h_out_g1); }
HG2 msm_finish_g2(const MsmPlan &p) { return msm_finish<HG2>(p, p.h_out_g2); }

// =====================================================================================================================
// proving key
static void upload_csr(const zkpk::Csr &h, DeviceCsr &d) {
    d.nnz = (uint32_t)h.col.size();
    ZK_CUDA(cudaMalloc(&d.rowptr, h.rowptr.size() * 4));
    ZK_CUDA(cudaMalloc(&d.col, (size_t)(d.nnz + 1) * 4));
    ZK_CUDA(cudaMalloc(&d.coef, (size_t)(d.nnz + 1) * 4));
    ZK_CUDA(cudaMemcpy(d.rowptr, h.rowptr.data(), h.rowptr.size() * 4, cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMemcpy(d.col, h.col.data(), (size_t)d.nnz * 4, cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMemcpy(d.coef, h.coef.data(), (size_t)d.nnz * 4, cudaMemcpyHostToDevice));
}
constexpr uint32_t MAX_WIDE = 64;
struct SeedDev { uint32_t base; uint32_t w[16]; uint32_t h[8]; };      // one SHA-256 compression gadget to expand on the GPU (see sha256_witness_kernel)
static_assert(sizeof(SeedDev) == 100, "seed layout is part of the C-ABI (zkw::CompressionSeed)");
constexpr uint32_t SHA_RUN = zkw::SHA_RUN_VARS, MAX_SEEDS = 32;
constexpr int MSM_C = 16;                  // window bits of the dense H-query MSM: 16 windows, 32768 buckets
// Witness queries (A, B, L): 97 % of the scalars are 0 or 1 ("ones" buckets), nearly all others fit 64 bits, so only a few thousand
// entries reach the windowed buckets.  With 16-bit windows they were scattered over 32768 buckets and the bucket reduction (a chain of
// ~35 dependent point additions per thread, 8192 threads each doing a 15-bit scalar multiplication) cost 0.3 ms and 14 % of the proof's
// multiply-pipe work for next to nothing; 8-bit windows leave 128 buckets.  The fixed-base tables double (32 windows) -- HBM is cheap.
constexpr int MSM_C_SIDE = 8;

template <class A> static A fetch_point(const void *dev, size_t idx) {
    A a; ZK_CUDA(cudaMemcpy(&a, (const char *)dev + idx * sizeof(A), sizeof(A), cudaMemcpyDeviceToHost)); return a;
}

struct LaneSync { std::mutex mu; std::condition_variable cv; };

static Lane *lane_create(const DevicePk *pk, int index) {
    Lane *ln = new Lane();
    ln->index = index;
    const size_t nw = pk->num_vars + 4, m = pk->dom->m;            // [1 | w | r | s | -rs]
    ZK_CUDA(cudaMalloc(&ln->w_can, nw * 32)); ZK_CUDA(cudaMalloc(&ln->w_mont, nw * 32));
    ZK_CUDA(cudaMemset(ln->w_can, 0, nw * 32));
    const uint64_t one_can[4] = {1, 0, 0, 0};
    ZK_CUDA(cudaMemcpy(ln->w_can, one_can, 32, cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMallocHost(&ln->h_w_pinned, (nw + 1) * 32));
    ZK_CUDA(cudaMalloc(&ln->w_lo, nw * 8)); ZK_CUDA(cudaMalloc(&ln->w_wide, MAX_WIDE * sizeof(WideIn)));
    ZK_CUDA(cudaMalloc(&ln->d_seeds, MAX_SEEDS * sizeof(SeedDev) + (MAX_SEEDS + 1) * 16));                  // seeds, then the segment table
    ZK_CUDA(cudaMallocHost(&ln->h_seeds_pinned, MAX_SEEDS * sizeof(SeedDev) + (MAX_SEEDS + 1) * 16));
    ZK_CUDA(cudaMallocHost(&ln->h_wide_pinned, MAX_WIDE * sizeof(WideIn)));
    ZK_CUDA(cudaMalloc(&ln->bufA, 3 * m * 32));                     // A | B | C contiguous: the three transforms of a stage share one launch
    ln->bufB = (char *)ln->bufA + m * 32; ln->bufC = (char *)ln->bufA + 2 * m * 32;
    ZK_CUDA(cudaMalloc(&ln->tmp, 3 * m * 32));
    ZK_CUDA(cudaMalloc(&ln->sat_flag, 4)); ZK_CUDA(cudaMallocHost(&ln->h_sat_flag, 4));
    // witness MSMs: ~97 % of the scalars are 0/1 and nearly all others are <= 64 bits; H is dense
    ln->mA.init(pk->nA, MSM_C_SIDE, 4096, true, false, true);
    ln->mB.init(pk->nB, MSM_C_SIDE, 4096, true, true, true);
    ln->mL.init(pk->nL, MSM_C_SIDE, 4096, true, false, true);
    // Batched-affine halving rounds in front of the H-query accumulation (msm.cuh): measured and left OFF -- 2 rounds cut the multiplications
    // per addition from 10 to 6 on paper, but with ~55 pairs per thread the two passes, the pair classification and the inversion put as
    // many instructions on the multiply pipe as the XYZZ additions they replace (profiles/r02_notes.md).  ZKB200_AFFINE_ROUNDS=2 enables them.
    int aff = 0;
    if (const char *e = getenv("ZKB200_AFFINE_ROUNDS")) aff = atoi(e);
    ln->mH.init(pk->nH, MSM_C, 0, true, false, true, aff);
    int prio_lo = 0, prio_hi = 0;
    ZK_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));          // the QAP map + H MSM chain is the critical path
    if (const char *e = getenv("ZKB200_SIDE_PRIO")) { if (atoi(e) == 1) { const int t = prio_lo; prio_lo = prio_hi; prio_hi = t; } else if (atoi(e) == 2) prio_lo = prio_hi; }
    ZK_CUDA(cudaStreamCreateWithPriority(&ln->s_main, cudaStreamNonBlocking, prio_hi));
    ZK_CUDA(cudaStreamCreateWithPriority(&ln->s_a, cudaStreamNonBlocking, prio_lo));
    ZK_CUDA(cudaStreamCreateWithPriority(&ln->s_b, cudaStreamNonBlocking, prio_lo));
    ZK_CUDA(cudaStreamCreateWithPriority(&ln->s_l, cudaStreamNonBlocking, prio_lo));
    ZK_CUDA(cudaStreamCreateWithPriority(&ln->s_b2, cudaStreamNonBlocking, prio_lo));
    ZK_CUDA(cudaStreamCreateWithPriority(&ln->s_sat, cudaStreamNonBlocking, prio_lo));
    ZK_CUDA(cudaEventCreateWithFlags(&ln->ev_w, cudaEventDisableTiming));
    ZK_CUDA(cudaMalloc(&ln->sig_dev, Lane::NSIG * 4)); ZK_CUDA(cudaMemset(ln->sig_dev, 0, Lane::NSIG * 4));
    ZK_CUDA(cudaMallocHost((void **)&ln->sig_host, Lane::NSIG * 4)); memset((void *)ln->sig_host, 0, Lane::NSIG * 4);
    cudaEvent_t *jev[] = {&ln->ev_ja, &ln->ev_jb, &ln->ev_jl, &ln->ev_jb2};
    for (auto *e : jev) ZK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    ZK_CUDA(cudaEventCreateWithFlags(&ln->ev_spmv, cudaEventDisableTiming));
    ZK_CUDA(cudaEventCreateWithFlags(&ln->ev_sat, cudaEventDisableTiming));
    cudaEvent_t *tev[] = {&ln->ev_t0, &ln->ev_t1, &ln->ev_q0, &ln->ev_q1, &ln->ev_h0, &ln->ev_h1, &ln->ev_a, &ln->ev_b, &ln->ev_l, &ln->ev_b2};
    for (auto *e : tev) ZK_CUDA(cudaEventCreate(e));
    return ln;
}
static void lane_destroy(Lane *ln) {
    void *ps[] = {ln->w_can, ln->w_mont, ln->bufA, ln->tmp, ln->sat_flag, ln->w_lo, ln->w_wide, ln->d_seeds};
    if (ln->h_seeds_pinned) cudaFreeHost(ln->h_seeds_pinned);
    for (void *p : ps) if (p) cudaFree(p);
    if (ln->h_wide_pinned) cudaFreeHost(ln->h_wide_pinned);
    if (ln->h_w_pinned) cudaFreeHost(ln->h_w_pinned);
    if (ln->h_sat_flag) cudaFreeHost(ln->h_sat_flag);
    ln->mA.release(); ln->mB.release(); ln->mH.release(); ln->mL.release();
    cudaStream_t ss[] = {ln->s_main, ln->s_a, ln->s_b, ln->s_l, ln->s_b2, ln->s_sat};
    for (auto st : ss) if (st) cudaStreamDestroy(st);
    cudaEvent_t es[] = {ln->ev_w, ln->ev_a, ln->ev_b, ln->ev_l, ln->ev_b2, ln->ev_t0, ln->ev_t1, ln->ev_q0, ln->ev_q1, ln->ev_h0, ln->ev_h1, ln->ev_spmv, ln->ev_sat, ln->ev_ja, ln->ev_jb, ln->ev_jl, ln->ev_jb2};
    for (auto &g : ln->graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ln->sig_dev) cudaFree(ln->sig_dev);
    if (ln->sig_host) cudaFreeHost((void *)ln->sig_host);
    for (auto e : es) if (e) cudaEventDestroy(e);
    delete ln;
}
Lane *lane_acquire(DevicePk *pk) {
    LaneSync *sy = (LaneSync *)pk->sync;
    std::unique_lock<std::mutex> lk(sy->mu);
    for (;;) {
        for (int i = 0; i < pk->nlanes; i++) if (!pk->lanes[i]->busy) { pk->lanes[i]->busy = true; return pk->lanes[i]; }
        sy->cv.wait(lk);
    }
}
Lane *lane_try(DevicePk *pk, int index) {
    if (index < 0 || index >= pk->nlanes) return nullptr;
    LaneSync *sy = (LaneSync *)pk->sync;
    std::lock_guard<std::mutex> lk(sy->mu);
    if (pk->lanes[index]->busy) return nullptr;
    pk->lanes[index]->busy = true;
    return pk->lanes[index];
}
void lane_release(DevicePk *pk, Lane *ln) {
    LaneSync *sy = (LaneSync *)pk->sync;
    { std::lock_guard<std::mutex> lk(sy->mu); ln->busy = false; }
    sy->cv.notify_one();
}
Lane *lane_of_staging(DevicePk *pk, const void *p) {
    for (int i = 0; i < pk->nlanes; i++) {
        const char *base = (const char *)pk->lanes[i]->h_w_pinned;
        if ((const char *)p >= base && (const char *)p < base + (pk->num_vars + 5) * 32) return pk->lanes[i];
    }
    return nullptr;
}

zkpk::ParsedPk *pk_parse_file(const char *path, std::string &err, double *seconds) {
    const double t0 = now_s();
    std::ifstream fh(path, std::ios::binary | std::ios::ate);
    if (!fh.is_open()) { err = std::string("cannot open proving key ") + path; return nullptr; }
    const size_t len = (size_t)fh.tellg();
    std::vector<uint8_t> data(len);
    fh.seekg(0); fh.read((char *)data.data(), (std::streamsize)len);
    zkpk::ParsedPk *P = new zkpk::ParsedPk();
    if (!zkpk::parse_pk(data.data(), len, *P)) { err = P->error; delete P; return nullptr; }
    // a B_query entry must be a whole (G2, G1) pair or absent: one skip flag serves both halves (kc_multiexp.tcc:21-89 never emits anything else)
    for (size_t i = 0; i < P->B_g1.size(); i++)
        if (((P->B_g1[i].flags ^ P->B_g2[i].flags) >> 1) & 1) { err = "proving key B_query entry " + std::to_string(i) + " has only one half at infinity"; delete P; return nullptr; }
    if (seconds) *seconds = now_s() - t0;
    return P;
}
void pk_parsed_free(zkpk::ParsedPk *P) { delete P; }

DevicePk *pk_load(const char *path, int device, std::string &err) {
    double parse_s = 0;
    zkpk::ParsedPk *P = pk_parse_file(path, err, &parse_s);
    if (!P) return nullptr;
    DevicePk *pk = pk_from_parsed(*P, device, err, parse_s);
    delete P;
    return pk;
}

DevicePk *pk_from_parsed(const zkpk::ParsedPk &P, int device, std::string &err, double parse_seconds) {
    const double t0 = now_s() - parse_seconds;
    device_init(device);
    DeviceExclusive excl;
    DevicePk *pk = new DevicePk();
    pk->device = device; pk->parse_seconds = parse_seconds;
    pk->num_inputs = P.num_inputs; pk->num_vars = P.num_inputs + P.num_aux; pk->num_constraints = P.num_constraints;
    if (P.A.size() != pk->num_vars + 1 || P.L.size() != pk->num_vars - pk->num_inputs || P.B_domain != pk->num_vars + 1) {
        err = "proving key query sizes do not match its constraint system"; delete pk; return nullptr;
    }
    pk->dom = Domain::build(pk->num_constraints + pk->num_inputs + 1);
    if (!pk->dom || P.H.size() != pk->dom->m - 1) { err = "proving key H_query does not match the evaluation domain"; delete pk; return nullptr; }

    // points: decompress on the GPU
    const SqrtConsts k = sqrt_consts();
    uint32_t *d_bad; ZK_CUDA(cudaMalloc(&d_bad, 4)); ZK_CUDA(cudaMemset(d_bad, 0, 4));
    const double t2 = now_s();
    decompress_vec<zkpk::CompressedG1, G1Affine>(P.A, (G1Affine **)&pk->A, &pk->A_skip, false, d_bad, k);
    decompress_vec<zkpk::CompressedG1, G1Affine>(P.B_g1, (G1Affine **)&pk->B1, &pk->B_skip, false, d_bad, k);
    decompress_vec<zkpk::CompressedG2, G2Affine>(P.B_g2, (G2Affine **)&pk->B2, nullptr, true, d_bad, k);
    decompress_vec<zkpk::CompressedG1, G1Affine>(P.H, (G1Affine **)&pk->H, &pk->H_skip, false, d_bad, k);
    decompress_vec<zkpk::CompressedG1, G1Affine>(P.L, (G1Affine **)&pk->L, &pk->L_skip, false, d_bad, k);
    std::vector<zkpk::CompressedG1> fixed1 = {P.alpha_g1, P.beta_g1, P.delta_g1};
    std::vector<zkpk::CompressedG2> fixed2 = {P.beta_g2, P.delta_g2};
    G1Affine *f1; G2Affine *f2;
    decompress_vec<zkpk::CompressedG1, G1Affine>(fixed1, &f1, nullptr, false, d_bad, k);
    decompress_vec<zkpk::CompressedG2, G2Affine>(fixed2, &f2, nullptr, true, d_bad, k);
    pk->alpha_g1 = fetch_point<HG1Affine>(f1, 0); pk->beta_g1 = fetch_point<HG1Affine>(f1, 1); pk->delta_g1 = fetch_point<HG1Affine>(f1, 2);
    pk->beta_g2 = fetch_point<HG2Affine>(f2, 0); pk->delta_g2 = fetch_point<HG2Affine>(f2, 1);
    uint32_t bad = 0; ZK_CUDA(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost)); cudaFree(d_bad);
    pk->decompress_seconds = now_s() - t2;
    if (bad) { err = "proving key holds " + std::to_string(bad) + " x-coordinates that are not on the curve"; pk_free(pk); return nullptr; }
    pk->nA = (uint32_t)P.A.size(); pk->nB = (uint32_t)P.B_g1.size(); pk->nH = (uint32_t)P.H.size(); pk->nL = (uint32_t)P.L.size();
    // The zero-knowledge terms r*delta_g1, s*(delta_g2, delta_g1), -(r*s)*delta_g1 ride along in the A, B, L MSMs: delta is appended as
    // one more base (decompress_vec left room), its scalar sits after the assignment in w_can = [1 | w | r | s | -rs].
    const uint32_t nw1 = (uint32_t)pk->num_vars + 1;                 // index of r in w_can; s = nw1 + 1; -rs = nw1 + 2
    const G1Affine d1 = fetch_point<G1Affine>(f1, 2); const G2Affine d2 = fetch_point<G2Affine>(f2, 1);
    const uint8_t zero8 = 0;
    ZK_CUDA(cudaMemcpy((G1Affine *)pk->A + pk->nA, &d1, sizeof d1, cudaMemcpyHostToDevice)); ZK_CUDA(cudaMemcpy(pk->A_skip + pk->nA, &zero8, 1, cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMemcpy((G1Affine *)pk->B1 + pk->nB, &d1, sizeof d1, cudaMemcpyHostToDevice)); ZK_CUDA(cudaMemcpy(pk->B_skip + pk->nB, &zero8, 1, cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMemcpy((G2Affine *)pk->B2 + pk->nB, &d2, sizeof d2, cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMemcpy((G1Affine *)pk->L + pk->nL, &d1, sizeof d1, cudaMemcpyHostToDevice)); ZK_CUDA(cudaMemcpy(pk->L_skip + pk->nL, &zero8, 1, cudaMemcpyHostToDevice));
    // a B entry whose G1 half is infinity but G2 half is not cannot be expressed with one skip flag; libsnark never emits one
    std::vector<uint32_t> bidx(P.B_idx); bidx.push_back(nw1 + 1);
    ZK_CUDA(cudaMalloc(&pk->B_idx, bidx.size() * 4));
    ZK_CUDA(cudaMemcpy(pk->B_idx, bidx.data(), bidx.size() * 4, cudaMemcpyHostToDevice));
    std::vector<uint32_t> lidx(pk->nL + 1);
    for (uint32_t i = 0; i < pk->nL; i++) lidx[i] = (uint32_t)pk->num_inputs + 1 + i;
    lidx[pk->nL] = nw1 + 2;
    ZK_CUDA(cudaMalloc(&pk->L_idx, lidx.size() * 4));
    ZK_CUDA(cudaMemcpy(pk->L_idx, lidx.data(), lidx.size() * 4, cudaMemcpyHostToDevice));
    pk->nA += 1; pk->nB += 1; pk->nL += 1;
    cudaFree(f1); cudaFree(f2);
    // fixed-base tables: 2^(16k) * P for every window k, so each MSM needs a single bucket set and no Horner step
    const double t3 = now_s();
    { void *e;
      e = msm_expand_bases(pk->A, pk->nA, MSM_C_SIDE, false); cudaFree(pk->A); pk->A = e;
      e = msm_expand_bases(pk->B1, pk->nB, MSM_C_SIDE, false); cudaFree(pk->B1); pk->B1 = e;
      e = msm_expand_bases(pk->B2, pk->nB, MSM_C_SIDE, true); cudaFree(pk->B2); pk->B2 = e;
      e = msm_expand_bases(pk->H, pk->nH, MSM_C, false); cudaFree(pk->H); pk->H = e;
      e = msm_expand_bases(pk->L, pk->nL, MSM_C_SIDE, false); cudaFree(pk->L); pk->L = e;
      device_sync(); }
    pk->expand_seconds = now_s() - t3;

    upload_csr(P.a, pk->a); upload_csr(P.b, pk->b); upload_csr(P.c, pk->c);
    pk->ncoef = (uint32_t)P.coef_dict.size();
    pk->coef_dict = dev_const(P.coef_dict.data(), P.coef_dict.size());

    int nl = 3;                                                    // proofs in flight per key
    if (const char *e = getenv("ZKB200_LANES")) nl = atoi(e);
    if (nl < 1) nl = 1;
    if (nl > MAX_LANES) nl = MAX_LANES;
    pk->sync = new LaneSync();
    for (int i = 0; i < nl; i++) { pk->lanes[i] = lane_create(pk, i); pk->nlanes = i + 1; }
    device_sync();
    pk->load_seconds = now_s() - t0;
    return pk;
}

uint64_t *compact_staging(Lane *ln) { return (uint64_t *)ln->h_w_pinned; }

void pk_free(DevicePk *pk) {
    if (!pk) return;
    cudaSetDevice(pk->device);
    DeviceExclusive excl;
    cudaDeviceSynchronize();
    void *ps[] = {pk->A, pk->B1, pk->B2, pk->H, pk->L, pk->A_skip, pk->B_skip, pk->H_skip, pk->L_skip, pk->B_idx, pk->L_idx, pk->a.rowptr, pk->a.col, pk->a.coef,
                  pk->b.rowptr, pk->b.col, pk->b.coef, pk->c.rowptr, pk->c.col, pk->c.coef, pk->coef_dict};
    for (void *p : ps) if (p) cudaFree(p);
    for (int i = 0; i < pk->nlanes; i++) lane_destroy(pk->lanes[i]);
    if (pk->dom) { pk->dom->release(); delete pk->dom; }
    delete (LaneSync *)pk->sync;
    delete pk;
}

// =====================================================================================================================
// the per-proof pipeline
static void upload_assignment(DevicePk *pk, Lane *ln, const uint8_t *assignment, const uint64_t *zk_scalars /* r, s, -rs or null */, cudaStream_t st) {
    const size_t n = pk->num_vars;
    char *pin = (char *)ln->h_w_pinned + 32;            // [pad | assignment | r | s | -rs]
    if (zk_scalars) memcpy(pin + n * 32, zk_scalars, 96);
    if (assignment) {
        if ((const char *)assignment != pin) memcpy(pin, assignment, n * 32);
        ZK_CUDA(cudaMemcpyAsync((char *)ln->w_can + 32, pin, n * 32 + (zk_scalars ? 96 : 0), cudaMemcpyHostToDevice, st));
        g_h2d_bytes += n * 32 + (zk_scalars ? 96 : 0);
        ZK_CUDA(cudaMemcpyAsync(ln->w_mont, ln->w_can, (n + 1) * 32, cudaMemcpyDeviceToDevice, st));
        ZK_LAUNCH(to_mont_kernel, cdiv(n + 1, 256), 256, 0, st, (Fr *)ln->w_mont, (uint32_t)(n + 1));
    } else if (zk_scalars) {                            // resident assignment: only r, s, -rs change
        ZK_CUDA(cudaMemcpyAsync((char *)ln->w_can + 32 + n * 32, pin + n * 32, 96, cudaMemcpyHostToDevice, st));
        g_h2d_bytes += 96;
    }
}

// compact upload: w_can[i] = lo[i] (zero-extended), w_mont[i] = lo[i] * R, then the few wide values (and r, s, -rs) are patched in
__global__ void expand_assignment_kernel(const uint64_t *__restrict__ lo, uint32_t count, Fr *__restrict__ w_can, Fr *__restrict__ w_mont) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t x = lo[i];
    Fr c = Fr::zero(); c.v[0] = (uint32_t)x; c.v[1] = (uint32_t)(x >> 32);
    st_fr(w_can + i, c);
    st_fr(w_mont + i, x <= 1 ? (x ? Fr::one() : c) : c.to_mont());
}
__global__ void patch_wide_kernel(const WideIn *__restrict__ wide, uint32_t nwide, uint32_t mont_limit, Fr *__restrict__ w_can, Fr *__restrict__ w_mont) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nwide) return;
    Fr c; memcpy(c.v, wide[k].v, 32);
    st_fr(w_can + wide[k].idx, c);
    if (wide[k].idx < mont_limit) st_fr(w_mont + wide[k].idx, c.to_mont());
}
// ---- SHA-256 gadget witness on the GPU -------------------------------------------------------------------------
// 97 % of a BlockMaze assignment are the internal variables of sha256_compression_function_gadget instances: per compression a contiguous
// run of 24 792 variables, every one of them a 32-bit word of the 64-round trace, a bit of such a word, or an unreduced sum with its
// overflow bits (layout: witness_sha.hpp).  The host generator (witness.cpp) hands over a 100-byte seed per compression -- the run's first
// variable, the 16 message words, the incoming chaining value -- and one CTA per compression re-runs the rounds (one thread, microseconds)
// and then writes the run straight into the canonical and the Montgomery assignment.
// SHA_PARTS CTAs share one compression: each rebuilds the trace (it is the serial part, ~15 us for one thread) and writes its slice of the run.
constexpr uint32_t SHA_PARTS = 8;
__global__ void __launch_bounds__(256) sha256_witness_kernel(const SeedDev *__restrict__ seeds, Fr *__restrict__ w_can, Fr *__restrict__ w_mont) {
    __shared__ zkw::ShaTrace T;
    const SeedDev sd = seeds[blockIdx.x / SHA_PARTS];
    if (threadIdx.x == 0) zkw::sha_trace_build(sd.w, sd.h, T);
    __syncthreads();
    const uint32_t per = (SHA_RUN + SHA_PARTS - 1) / SHA_PARTS, v0 = (blockIdx.x % SHA_PARTS) * per, v1 = min(v0 + per, SHA_RUN);
    for (uint32_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        const uint64_t x = zkw::sha_trace_value(T, v);
        Fr c = Fr::zero(); c.v[0] = (uint32_t)x; c.v[1] = (uint32_t)(x >> 32);
        st_fr(w_can + sd.base + v, c);
        st_fr(w_mont + sd.base + v, x <= 1 ? (x ? Fr::one() : c) : c.to_mont());
    }
}

static void upload_compact(DevicePk *pk, Lane *ln, const uint64_t *lo, const WideIn *wide, uint32_t nwide, const uint64_t *zk_scalars, cudaStream_t st) {
    const uint32_t n = (uint32_t)pk->num_vars;
    uint64_t *pin = (uint64_t *)ln->h_w_pinned;
    if (lo != pin) memcpy(pin, lo, (size_t)(n + 1) * 8);
    WideIn *pw = (WideIn *)ln->h_wide_pinned;
    if (nwide) memcpy(pw, wide, nwide * sizeof(WideIn));             // nwide <= MAX_WIDE - 3 and every idx in [1, n]: checked by prove_submit
    for (int k = 0; k < 3; k++) { pw[nwide + k].idx = n + 1 + k; pw[nwide + k].pad = 0; memcpy(pw[nwide + k].v, zk_scalars + 4 * k, 32); }
    nwide += 3;
    ZK_CUDA(cudaMemcpyAsync(ln->w_lo, pin, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    ZK_CUDA(cudaMemcpyAsync(ln->w_wide, pw, nwide * sizeof(WideIn), cudaMemcpyHostToDevice, st));
    g_h2d_bytes += (size_t)(n + 1) * 8 + nwide * sizeof(WideIn);
    ZK_LAUNCH(expand_assignment_kernel, cdiv(n + 1, 256), 256, 0, st, (const uint64_t *)ln->w_lo, n + 1, (Fr *)ln->w_can, (Fr *)ln->w_mont);
    ZK_LAUNCH(patch_wide_kernel, 1, 64, 0, st, (const WideIn *)ln->w_wide, nwide, n + 1, (Fr *)ln->w_can, (Fr *)ln->w_mont);
}
// With seeds, only the variables OUTSIDE the compression runs travel: they are packed to the front of the pinned staging buffer (a few
// dozen KB instead of 1.8 MB for send) together with a table of (first variable, count) segments, expanded by one small kernel, and the
// runs are written by sha256_witness_kernel.
struct SegDev { uint32_t dst, len, src, pad; };
constexpr uint32_t MAX_SEGS = MAX_SEEDS + 1;
__global__ void expand_segments_kernel(const uint64_t *__restrict__ packed, const SegDev *__restrict__ segs, uint32_t nseg, uint32_t total,
                                       Fr *__restrict__ w_can, Fr *__restrict__ w_mont) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t k = 0;
    while (k + 1 < nseg && i >= segs[k + 1].src) k++;
    const uint32_t dst = segs[k].dst + (i - segs[k].src);
    const uint64_t x = packed[i];
    Fr c = Fr::zero(); c.v[0] = (uint32_t)x; c.v[1] = (uint32_t)(x >> 32);
    st_fr(w_can + dst, c);
    st_fr(w_mont + dst, x <= 1 ? (x ? Fr::one() : c) : c.to_mont());
}
static void upload_compact_seeded(DevicePk *pk, Lane *ln, const uint64_t *lo, const WideIn *wide, uint32_t nwide, const SeedDev *seeds, uint32_t nseeds,
                                  const uint64_t *zk_scalars, cudaStream_t st) {
    const uint32_t n = (uint32_t)pk->num_vars;
    // runs in increasing order of their first variable (the generators emit them that way; sort anyway: the input is untrusted)
    SeedDev *hs = (SeedDev *)ln->h_seeds_pinned;
    memcpy(hs, seeds, (size_t)nseeds * sizeof(SeedDev));
    std::sort(hs, hs + nseeds, [](const SeedDev &a, const SeedDev &b) { return a.base < b.base; });
    SegDev *segs = (SegDev *)((char *)ln->h_seeds_pinned + MAX_SEEDS * sizeof(SeedDev));
    uint64_t *pin = (uint64_t *)ln->h_w_pinned;
    uint32_t nseg = 0, packed = 0, cur = 0;                      // cur: next variable not yet covered
    for (uint32_t k = 0; k <= nseeds; k++) {
        const uint32_t stop = k < nseeds ? hs[k].base : n + 1;
        if (stop > cur) {
            segs[nseg++] = SegDev{cur, stop - cur, packed, 0};
            memmove(pin + packed, lo + cur, (size_t)(stop - cur) * 8);      // in place when lo is this lane's staging buffer (packed <= cur)
            packed += stop - cur;
        }
        if (k < nseeds) cur = hs[k].base + SHA_RUN > cur ? hs[k].base + SHA_RUN : cur;      // overlapping runs: still every variable covered once
    }
    WideIn *pw = (WideIn *)ln->h_wide_pinned;
    if (nwide) memcpy(pw, wide, nwide * sizeof(WideIn));
    for (int k = 0; k < 3; k++) { pw[nwide + k].idx = n + 1 + k; pw[nwide + k].pad = 0; memcpy(pw[nwide + k].v, zk_scalars + 4 * k, 32); }
    nwide += 3;
    ZK_CUDA(cudaMemcpyAsync(ln->w_lo, pin, (size_t)packed * 8, cudaMemcpyHostToDevice, st));
    ZK_CUDA(cudaMemcpyAsync(ln->d_seeds, hs, MAX_SEEDS * sizeof(SeedDev) + (size_t)nseg * sizeof(SegDev), cudaMemcpyHostToDevice, st));
    ZK_CUDA(cudaMemcpyAsync(ln->w_wide, pw, nwide * sizeof(WideIn), cudaMemcpyHostToDevice, st));
    g_h2d_bytes += (size_t)packed * 8 + MAX_SEEDS * sizeof(SeedDev) + (size_t)nseg * sizeof(SegDev) + nwide * sizeof(WideIn);
    const SegDev *dsegs = (const SegDev *)((const char *)ln->d_seeds + MAX_SEEDS * sizeof(SeedDev));
    if (packed) ZK_LAUNCH(expand_segments_kernel, cdiv(packed, 256), 256, 0, st, (const uint64_t *)ln->w_lo, dsegs, nseg, packed, (Fr *)ln->w_can, (Fr *)ln->w_mont);
    ZK_LAUNCH(sha256_witness_kernel, nseeds * SHA_PARTS, 256, 0, st, (const SeedDev *)ln->d_seeds, (Fr *)ln->w_can, (Fr *)ln->w_mont);
    ZK_LAUNCH(patch_wide_kernel, 1, 64, 0, st, (const WideIn *)ln->w_wide, nwide, n + 1, (Fr *)ln->w_can, (Fr *)ln->w_mont);
}

// r1cs_to_qap_witness_map (r1cs_to_qap.tcc:205-334) on stream st.  Result: coefficients_for_H[0..m) in ln->tmp (Montgomery form).
static void qap_pipeline(DevicePk *pk, Lane *ln, cudaStream_t st) {
    const Domain &d = *pk->dom;
    const uint32_t nc = (uint32_t)pk->num_constraints, m = d.m;
    const Fr *w = (const Fr *)ln->w_mont, *dict = (const Fr *)pk->coef_dict;
    Fr *A = (Fr *)ln->bufA, *B = (Fr *)ln->bufB, *C = (Fr *)ln->bufC, *T = (Fr *)ln->tmp;
    SpmvArgs sa{{pk->a.rowptr, pk->b.rowptr, pk->c.rowptr}, {pk->a.col, pk->b.col, pk->c.col}, {pk->a.coef, pk->b.coef, pk->c.coef}, {A, B, C}};
    ZK_LAUNCH(spmv_kernel, dim3(cdiv((size_t)m * SPMV_G, 128), 3), 128, 0, st, sa, dict, w, nc, m, (uint32_t)pk->num_inputs, g_spmv_bits);
    // is_satisfied only reads the evaluation vectors, which stay intact until the forward transforms write them: it runs on its own stream
    // beside the inverse transforms instead of in front of them
    ZK_CUDA(cudaEventRecord(ln->ev_spmv, st));
    ZK_CUDA(cudaStreamWaitEvent(ln->s_sat, ln->ev_spmv, 0));
    ZK_CUDA(cudaMemsetAsync(ln->sat_flag, 0, 4, ln->s_sat));
    ZK_LAUNCH(sat_check_kernel, cdiv(nc, 256), 256, 0, ln->s_sat, A, B, C, nc, ln->sat_flag);
    ZK_CUDA(cudaMemcpyAsync(ln->h_sat_flag, ln->sat_flag, 4, cudaMemcpyDeviceToHost, ln->s_sat));
    ZK_CUDA(cudaEventRecord(ln->ev_sat, ln->s_sat));
    // iFFT then cosetFFT of each of A, B, C: coefficient i is multiplied by g^i (and 1/m for the basic domain) on the way
    domain_ifft(st, d, A, T, pm_none(), pm_two(d.g_lo, d.g_hi), 3, m);
    ZK_CUDA(cudaStreamWaitEvent(st, ln->ev_sat, 0));
    if (!d.step && g_qap_skip_c) {
        // Basic domain, six transforms instead of the reference's seven (r1cs_to_qap.tcc:240-311 also takes c to the coset).  With
        // Z = x^m - 1:  a*b = c + Z*h, deg h <= m-2, and on the coset g*S the polynomial x^m is the constant g^m, so the interpolant of the
        // coset values of a*b is  d = a*b mod (x^m - g^m) = c + (g^m - 1) h,  i.e.  h_i = (d_i - c_i) / (g^m - 1)  coefficient by
        // coefficient -- c never has to be evaluated on the coset.  T holds m*a | m*b | m*c (the inverse transforms leave the 1/m to the
        // next multiplication), so the last pass of the final inverse transform computes  raw_i * g^-i/(m Z) - (m c_i) * 1/(m Z).
        // Valid for a satisfying assignment (c interpolates a.*b on S); otherwise the default proof is returned anyway.
        domain_fft(st, d, T, A, d.g_full_ninv ? pm_full(d.g_full_ninv) : pm_two(d.g_lo, d.g_hi_ninv), 2, m);
        ZK_LAUNCH(qap_product_kernel, cdiv(m, 256), 256, 0, st, A, (const Fr *)B, m);
        ntt(st, A, T, d.tw_big_i, d.log_big, pm_none(), d.gi_full_zninv ? pm_full(d.gi_full_zninv) : pm_two(d.gi_lo, d.gi_hi_zninv), 1, 0, T + 2 * (size_t)m, d.c_z_over_m);
        return;
    }
    if (!d.step) domain_fft(st, d, T, A, d.g_full_ninv ? pm_full(d.g_full_ninv) : pm_two(d.g_lo, d.g_hi_ninv), 3, m);
    else domain_fft(st, d, T, A, pm_none(), 3, m);
    ZK_LAUNCH(qap_pointwise_kernel, cdiv(m, 256), 256, 0, st, A, B, C, m, d.big, d.compr, (const Fr *)d.zt, to_dev(d.z1));
    domain_ifft(st, d, A, T, pm_two(d.gi_lo, d.gi_hi_ninv), pm_two(d.gi_lo, d.gi_hi));
}

int qap_witness_map(DevicePk *pk, const uint8_t *assignment, uint8_t *out_H, int *satisfied) {
    device_init(pk->device);
    Lane *ln = lane_acquire(pk);
    cudaStream_t st = ln->s_main;
    upload_assignment(pk, ln, assignment, nullptr, st);
    qap_pipeline(pk, ln, st);
    const uint32_t m = pk->dom->m;
    ZK_LAUNCH(from_mont_kernel, cdiv(m, 256), 256, 0, st, (const Fr *)ln->tmp, (Fr *)ln->bufB, m);
    ZK_CUDA(cudaMemcpyAsync(out_H, ln->bufB, (size_t)m * 32, cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    memset(out_H + (size_t)m * 32, 0, 32);          // coefficients_for_H[m] = 0 (no ZK patch: d1 = d2 = d3 = 0)
    if (satisfied) *satisfied = (*ln->h_sat_flag == 0);
    lane_release(pk, ln);
    return 0;
}

// parity hook: the canonical assignment [w_1 .. w_n] a lane holds (what the last proof on it was made for)
int lane_read_assignment(DevicePk *pk, Lane *ln, uint8_t *out) {
    device_init(pk->device);
    ZK_CUDA(cudaStreamSynchronize(ln->s_main));
    ZK_CUDA(cudaMemcpy(out, (const char *)ln->w_can + 32, (size_t)pk->num_vars * 32, cudaMemcpyDeviceToHost));
    return 0;
}

static HG1 g1_mul(const HG1Affine &p, const uint64_t k[4]) { return HG1::from_affine(p).mul(k); }

int prove_submit(DevicePk *pk, Lane *ln, const uint8_t *assignment, const uint64_t *lo, const WideIn *wide, uint32_t nwide, const uint64_t r[4],
                 const uint64_t s[4], const void *seeds, uint32_t nseeds) {
    // malformed input is an error, never a proof of something else
    if (HFr::geq_mod(r) || HFr::geq_mod(s)) return -4;
    if (lo) {
        if (nwide > MAX_WIDE - 3 || (nwide && !wide)) return -2;
        for (uint32_t k = 0; k < nwide; k++) if (wide[k].idx == 0 || wide[k].idx > pk->num_vars) return -3;
        if (nseeds > MAX_SEEDS || (nseeds && !seeds)) return -5;
        for (uint32_t k = 0; k < nseeds; k++) {
            const uint32_t base = ((const SeedDev *)seeds)[k].base;
            if (base == 0 || (uint64_t)base + SHA_RUN > pk->num_vars + 1) return -5;
        }
    } else if (nseeds) return -5;
    device_init(pk->device);
    g_launches = 0; g_h2d_bytes = 0;
    cudaStream_t st = ln->s_main;
    memcpy(ln->r, r, 32); memcpy(ln->s, s, 32);
    uint64_t zks[12];
    memcpy(zks, r, 32); memcpy(zks + 4, s, 32);
    (HFr::from_canonical(r) * HFr::from_canonical(s)).neg().to_canonical(zks + 8);          // -(r*s) mod r
    ZK_CUDA(cudaEventRecord(ln->ev_t0, st));
    if (lo && nseeds) upload_compact_seeded(pk, ln, lo, wide, nwide, (const SeedDev *)seeds, nseeds, zks, st);
    else if (lo) upload_compact(pk, ln, lo, wide, nwide, zks, st);
    else upload_assignment(pk, ln, assignment, zks, st);
    const int upload_launches = g_launches;
    ln->h2d_bytes = g_h2d_bytes;
    // no other proof in flight on this device (a benign race: only the schedule of the MSMs depends on it).  Device-wide: a mixed batch keeps
    // proofs of OTHER keys in flight on the same GPU, and the latency schedule -- team tails, two waves -- costs throughput there just the same.
    bool alone = true;
    for (int i = 0; i < pk->nlanes; i++) if (pk->lanes[i] != ln && pk->lanes[i]->pending) alone = false;
    if (pk->device >= 0 && pk->device < 64 && g_device_pending[pk->device].fetch_add(1, std::memory_order_relaxed) > 0) alone = false;
    ln->mA.latency = ln->mB.latency = ln->mL.latency = ln->mH.latency = alone;
    ln->mH.alone = alone;
    // Everything behind the upload is the same sequence of launches for every proof on this lane (given `alone`): it is captured into a
    // CUDA graph the first time and replayed with ONE driver call afterwards.  The ~90 driver calls a proof otherwise takes are nothing for
    // one GPU, but the driver serialises them per process: one process feeding 8 GPUs topped out at 3 080 proofs/s where 8 processes reached
    // 4 800 (profiles/r02_notes.md).  ZKB200_GRAPH=0 enqueues kernel by kernel as before; the roofline's isolate mode always does.
    const bool use_graph = g_use_graph && !g_isolate_h;
    Lane::LaneGraph &lg = ln->graph[alone ? 1 : 0];
    auto signal = [&](int k, cudaStream_t on) { ZK_LAUNCH(signal_kernel, 1, 1, 0, on, ln->sig_host + k, ln->sig_dev + k); };
    if (use_graph && lg.exec) {
        msm_set_mode(ln->mA); msm_set_mode(ln->mB); msm_set_mode(ln->mL); msm_set_mode(ln->mH);
        ZK_CUDA(cudaGraphLaunch(lg.exec, st));
        ZK_CUDA(cudaEventRecord(ln->ev_t1, st));
        ln->launches = upload_launches + lg.launches;
        ln->sig_epoch++;
        ln->pending = true;
        return 0;
    }
    std::shared_lock<std::shared_mutex> capture_lock(g_capture_mu[pk->device & 63], std::defer_lock);
    if (use_graph) { capture_lock.lock(); ZK_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed)); g_capturing = true; }
    ZK_CUDA(cudaEventRecord(ln->ev_w, st));
    // A, B, L queries on side streams: scalars are the canonical padded assignment [1 | w]  (r1cs_gg_ppzksnark.tcc:437-484); the trailing
    // delta base of each query picks up r, s, -rs, so the MSM results already are  eA + r*delta,  eB + s*delta,  eL - rs*delta.
    // (Starting them after the QAP map instead -- beside the H query -- was measured for a proof alone on the GPU: the QAP map drops from
    // 0.65 to 0.51 ms and the H query grows by as much; even a single proof keeps the multiply pipe busy, profiles/r02_notes.md.)
    auto side_queries = [&]() {
        ZK_CUDA(cudaStreamWaitEvent(ln->s_a, ln->ev_w, 0));
        ZK_CUDA(cudaStreamWaitEvent(ln->s_b, ln->ev_w, 0));
        ZK_CUDA(cudaStreamWaitEvent(ln->s_l, ln->ev_w, 0));
        msm_run(ln->s_b, ln->mB, ScalarRef{ln->w_can, pk->B_idx, 0, 0}, pk->B_skip, pk->B1, pk->B2, ln->s_b2);
        msm_run(ln->s_a, ln->mA, ScalarRef{ln->w_can, nullptr, 0, 0}, pk->A_skip, pk->A, nullptr);
        msm_run(ln->s_l, ln->mL, ScalarRef{ln->w_can, pk->L_idx, 0, 0}, pk->L_skip, pk->L, nullptr);
        // signals: what prove_collect polls; ev_a .. ev_b2: when the queries finished (statistics); ev_ja .. ev_jb2: the joins into s_main
        signal(Lane::SIG_A, ln->s_a); signal(Lane::SIG_B, ln->s_b); signal(Lane::SIG_L, ln->s_l); signal(Lane::SIG_B2, ln->s_b2);
        record_for_host(ln->ev_a, ln->s_a); record_for_host(ln->ev_b, ln->s_b); record_for_host(ln->ev_l, ln->s_l); record_for_host(ln->ev_b2, ln->s_b2);
        ZK_CUDA(cudaEventRecord(ln->ev_ja, ln->s_a)); ZK_CUDA(cudaEventRecord(ln->ev_jb, ln->s_b)); ZK_CUDA(cudaEventRecord(ln->ev_jl, ln->s_l));
        ZK_CUDA(cudaEventRecord(ln->ev_jb2, ln->s_b2));
    };
    auto wait_side = [&]() {
        ZK_CUDA(cudaStreamWaitEvent(st, ln->ev_ja, 0)); ZK_CUDA(cudaStreamWaitEvent(st, ln->ev_jb, 0)); ZK_CUDA(cudaStreamWaitEvent(st, ln->ev_jl, 0));
        ZK_CUDA(cudaStreamWaitEvent(st, ln->ev_jb2, 0));
    };
    // H: QAP witness map, then the dense MSM over coefficients_for_H[0 .. m-1).  This is the critical path, so the QAP map is enqueued
    // first (the host needs ~0.2 ms to enqueue the side queries).  Measurement mode: H waits for the side queries.
    record_for_host(ln->ev_q0, st);
    qap_pipeline(pk, ln, st);
    record_for_host(ln->ev_q1, st);
    side_queries();                           // enqueued while the GPU is busy with the QAP map: they run beside it and are mostly done when H starts
    if (g_isolate_h) wait_side();
    record_for_host(ln->ev_h0, st);
    msm_run(st, ln->mH, ScalarRef{ln->tmp, nullptr, 0, 1}, pk->H_skip, pk->H, nullptr);
    record_for_host(ln->ev_h1, st);
    wait_side();
    signal(Lane::SIG_DONE, st);
    if (use_graph) {
        g_capturing = false;
        cudaGraph_t graph = nullptr;
        ZK_CUDA(cudaStreamEndCapture(st, &graph));
        capture_lock.unlock();
        ZK_CUDA(cudaGraphInstantiate(&lg.exec, graph, 0));
        ZK_CUDA(cudaGraphDestroy(graph));
        lg.launches = g_launches - upload_launches;
        ZK_CUDA(cudaGraphLaunch(lg.exec, st));
    }
    ZK_CUDA(cudaEventRecord(ln->ev_t1, st));            // plain record behind the proof: the end of its GPU time
    ln->sig_epoch++;
    ln->launches = g_launches;
    ln->pending = true;
    return 0;
}

// Waiting for the GPU.  cudaEventSynchronize / cudaStreamSynchronize hand the thread to the driver, whose wake-up costs 0.1-0.3 ms on
// some hosts -- a tenth of a proof.  The collecting thread polls the event instead (ZKB200_SPIN=0 goes back to the blocking calls).
static const bool g_spin = [] { const char *e = getenv("ZKB200_SPIN"); return !(e && atoi(e) == 0); }();
static void wait_event(cudaEvent_t ev) {
    if (!g_spin) { ZK_CUDA(cudaEventSynchronize(ev)); return; }
    for (;;) {
        const cudaError_t e = cudaEventQuery(ev);
        if (e == cudaSuccess) return;
        if (e != cudaErrorNotReady) ZK_CUDA(e);
#if defined(__x86_64__)
        for (int i = 0; i < 32; i++) __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
}

// A faulted GPU never sets the flag: every few milliseconds of waiting the event recorded behind the proof is queried, which reports the
// sticky error of a crashed kernel -- and the library aborts loudly instead of hanging.
static void wait_signal(const Lane *ln, int k) {
    const volatile uint32_t *flag = ln->sig_host + k;
    const uint32_t want = ln->sig_epoch;
    auto alive = [&]() { const cudaError_t e = cudaEventQuery(ln->ev_t1); if (e != cudaSuccess && e != cudaErrorNotReady) ZK_CUDA(e); };
    if (!g_spin) {
        for (uint32_t n = 1; (int32_t)(*flag - want) < 0; n++) { std::this_thread::sleep_for(std::chrono::microseconds(20)); if ((n & 255u) == 0) alive(); }
        return;
    }
    for (uint32_t n = 1; (int32_t)(*flag - want) < 0; n++) {
#if defined(__x86_64__)
        for (int i = 0; i < 32; i++) __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
        if ((n & 8191u) == 0) alive();
    }
}

int prove_collect(DevicePk *pk, Lane *ln, ProofPoints &out) {
    if (!ln->pending) return -1;
    device_init(pk->device);
    const uint64_t *r = ln->r, *s = ln->s;
    // The witness queries finish well before the H query (which waits for the QAP map): everything of the proof combination
    // (r1cs_gg_ppzksnark.tcc:487-495) that does not involve H -- above all the two 254-bit scalar multiplications s*A and r*B1, 0.2 ms of
    // host time -- is done while the GPU is still busy with it.
    // The G1 queries first: the G2 half of B (3x the field work, the longest bucket reduction) is the last side chain to finish -- for the
    // smaller circuits barely before the H query -- and nothing below needs it until out.B.
    wait_signal(ln, Lane::SIG_A); wait_signal(ln, Lane::SIG_B); wait_signal(ln, Lane::SIG_L);
    const HG1 eAr = msm_finish_g1(ln->mA), eB1s = msm_finish_g1(ln->mB), eLrs = msm_finish_g1(ln->mL);
    const HG1 gA = HG1::from_affine(pk->alpha_g1).add(eAr);
    const HG1 g1B = HG1::from_affine(pk->beta_g1).add(eB1s);
    const HG1 c_part = eLrs.add(gA.mul(s)).add(g1B.mul(r));
    out.A = gA.to_affine();
    wait_signal(ln, Lane::SIG_B2);
    const HG2 eB2s = msm_finish_g2(ln->mB);
    const HG2 g2B = HG2::from_affine(pk->beta_g2).add(eB2s);
    out.B = g2B.to_affine();
    if (out.want_parts) {
        // the plain MSM values of the reference (parity hooks): strip the folded zero-knowledge terms again
        const HG1 rd = g1_mul(pk->delta_g1, r).neg(), sd = g1_mul(pk->delta_g1, s).neg();
        uint64_t rs[4]; (HFr::from_canonical(r) * HFr::from_canonical(s)).to_canonical(rs);
        out.At = eAr.add(rd).to_affine(); out.Bt_h = eB1s.add(sd).to_affine();
        out.Lt = eLrs.add(g1_mul(pk->delta_g1, rs)).to_affine();
        out.Bt_g = eB2s.add(HG2::from_affine(pk->delta_g2).mul(s).neg()).to_affine();
    }

    wait_signal(ln, Lane::SIG_DONE);                // the last thing on s_main, after the side streams joined
    wait_event(ln->ev_t1);                          // (recorded right behind it: the timings below need it complete)
    const double t_sync = now_s();
    ln->pending = false;
    if (pk->device >= 0 && pk->device < 64) g_device_pending[pk->device].fetch_sub(1, std::memory_order_relaxed);
    ZK_CUDA(cudaEventElapsedTime(&out.gpu_ms, ln->ev_t0, ln->ev_t1));
    out.qap_ms = elapsed_ms(ln->ev_q0, ln->ev_q1);
    out.msm_h_ms = elapsed_ms(ln->ev_h0, ln->ev_h1);
    out.acc_h_ms = ln->mH.last_acc_ms();
    out.a_done_ms = elapsed_ms(ln->ev_t0, ln->ev_a);
    { const float b1 = elapsed_ms(ln->ev_t0, ln->ev_b), b2 = elapsed_ms(ln->ev_t0, ln->ev_b2); out.b_done_ms = b1 > b2 ? b1 : b2; }
    out.l_done_ms = elapsed_ms(ln->ev_t0, ln->ev_l);
    out.satisfied = (*ln->h_sat_flag == 0);
    out.launches = ln->launches;
    g_last_launches = ln->launches;
    // what crossed PCIe for this proof: the uploads of prove_submit; back: the partial sums of the four queries and the is_satisfied flag
    auto d2h = [](const MsmPlan &p) {
        const unsigned long long per = (p.team_now & 4) ? ((p.team_dbg & 1) ? p.bpw_team : 1u) : p.bpw;
        return (unsigned long long)(p.regions + 1) * per * ((p.out_g1 ? sizeof(G1XYZZ) : 0) + (p.out_g2 ? sizeof(G2XYZZ) : 0));
    };
    g_last_bytes[0] = ln->h2d_bytes; g_last_bytes[1] = 4 + d2h(ln->mA) + d2h(ln->mB) + d2h(ln->mL) + d2h(ln->mH);
    const HG1 eH = msm_finish_g1(ln->mH);
    if (out.want_parts) out.Ht = eH.to_affine();
    out.C = eH.add(c_part).to_affine();
    out.host_tail_ms = (float)(1e3 * (now_s() - t_sync));
    return 0;
}

int prove(DevicePk *pk, const uint8_t *assignment, const uint64_t r[4], const uint64_t s[4], ProofPoints &out) {
    // a null assignment reuses what lane 0 holds (single-lane callers: the parity tests and the bench's sequential leg)
    Lane *ln = nullptr;
    if (!assignment) { while (!(ln = lane_try(pk, 0))) std::this_thread::yield(); }
    else ln = lane_acquire(pk);
    int rc = prove_submit(pk, ln, assignment, nullptr, nullptr, 0, r, s);
    if (rc == 0) rc = prove_collect(pk, ln, out);
    lane_release(pk, ln);
    return rc;
}
int prove_compact(DevicePk *pk, const uint64_t *lo, const WideIn *wide, uint32_t nwide, const uint64_t r[4], const uint64_t s[4], ProofPoints &out,
                  const void *seeds, uint32_t nseeds) {
    Lane *own = lane_of_staging(pk, lo);               // the caller generated the witness in a lane it already holds
    Lane *ln = own ? own : lane_acquire(pk);
    int rc = prove_submit(pk, ln, nullptr, lo, wide, nwide, r, s, seeds, nseeds);
    if (rc == 0) rc = prove_collect(pk, ln, out);
    if (!own) lane_release(pk, ln);
    return rc;
}

static void hex_fq(const HFq &x, std::string &o) {
    uint64_t c[4]; x.to_canonical(c);
    char buf[65];
    snprintf(buf, sizeof buf, "%016llx%016llx%016llx%016llx", (unsigned long long)c[3], (unsigned long long)c[2], (unsigned long long)c[1],
             (unsigned long long)c[0]);
    o += buf;
}
std::string proof_hex(const ProofPoints &p) {
    // A.x A.y B.x.c1 B.x.c0 B.y.c1 B.y.c0 C.x C.y (mintcgo.cpp:130-187); infinity prints as the affine form of (0,1,0)
    std::string o; o.reserve(512);
    auto g1 = [&](const HG1Affine &a) { if (a.is_inf()) { hex_fq(HFq::zero(), o); hex_fq(HFq::one(), o); } else { hex_fq(a.x, o); hex_fq(a.y, o); } };
    g1(p.A);
    if (p.B.is_inf()) { hex_fq(HFq::zero(), o); hex_fq(HFq::zero(), o); hex_fq(HFq::zero(), o); hex_fq(HFq::one(), o); }
    else { hex_fq(p.B.x.c1, o); hex_fq(p.B.x.c0, o); hex_fq(p.B.y.c1, o); hex_fq(p.B.y.c0, o); }
    g1(p.C);
    return o;
}

} // namespace zkp
