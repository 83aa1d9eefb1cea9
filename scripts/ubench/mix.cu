// Do the FP64 pipe and the integer-multiply pipe of a B200 SM run modular multiplications SIDE BY SIDE?  
//   role A: ff.cuh Montgomery multiplication, 8 x 32-bit limbs, IMAD.WIDE.U32.X carry chains       (multiply pipe)
//   role B: f52_mul below, 5 x 52-bit limbs held in doubles, exact hi/lo halves by two fma.rz + one add,
//           columns summed as 64-bit integers taken from the bit patterns, lazy (result < 2q, R = 2^260)  (FP64 + ALU pipes)
// CTAs of 128 threads; a CTA's role follows blockIdx.x / SMs, so each SM sub-partition holds warps of both roles.
// Host mode (no GPU): ./mix host N  prints N products of the same f52_mul for scripts/proto/fp64_mont_check.py --lazy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I blockmaze_b200/csrc scripts/ubench/mix.cu -o scripts/ubench/mix
#include <cfenv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include <cuda_runtime.h>
#include "ff.cuh"
using namespace zk;

struct F52 { double l[5]; };
#define F52_M52 ((1ull << 52) - 1)
#define F52_BH 0x4670000000000000ull          // bit pattern of 2^104: every high half carries it
#define F52_BL 0x4330000000000000ull          // bit pattern of 2^52: every low half carries it
#define F52_NINV 0x9f72fcf6d4d8fULL           // placeholder, replaced below by the constant computed at start-up (host) / passed in (device)

__host__ __device__ __forceinline__ double f52_fma_rz(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rz(a, b, c);
#else
    return std::fma(a, b, c);                 // host: the caller runs under fesetround(FE_TOWARDZERO)
#endif
}
__host__ __device__ __forceinline__ unsigned long long f52_bits(double d) {
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(d);
#else
    unsigned long long r; memcpy(&r, &d, 8); return r;
#endif
}
__host__ __device__ __forceinline__ double f52_from_bits(unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
// integer < 2^52 -> double
__host__ __device__ __forceinline__ double f52_from_int(unsigned long long x) { return f52_from_bits(x | F52_BL) - 0x1p52; }

// (a * b) / 2^260 mod q up to a multiple of q: inputs < 8q with limbs < 2^52, output < 2q with limbs < 2^52.
__host__ __device__ __forceinline__ F52 f52_mul(const F52 &a, const F52 &b, const unsigned long long ninv) {
    const double q[5] = {(double)0x08c16d87cfd47ull, (double)0x916871ca8d3c2ull, (double)0x181585d97816aull, (double)0xa029b85045b68ull, (double)0x30644e72e131ull};
    unsigned long long c[11];
    // every column starts at minus the biases it will ever receive (multiples of 2^52: the low 52 bits, which the reduction reads before
    // a column is complete, do not see them)
#pragma unroll
    for (int k = 0; k < 11; k++) {
        const int nl_ab = k < 5 ? k + 1 : (k < 9 ? 9 - k : 0), nh_ab = k == 0 ? 0 : (k <= 5 ? k : (k <= 9 ? 10 - k : 0));
        const int nl_mq = nl_ab, nh_mq = nh_ab;                              // the m*q products have the same shape
        c[k] = 0ull - ((unsigned long long)(nl_ab + nl_mq) * F52_BL + (unsigned long long)(nh_ab + nh_mq) * F52_BH);
    }
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const double h = f52_fma_rz(a.l[i], b.l[j], 0x1p104);
            const double l = f52_fma_rz(a.l[i], b.l[j], (0x1p104 + 0x1p52) - h);
            c[i + j] += f52_bits(l); c[i + j + 1] += f52_bits(h);
        }
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const unsigned long long t = c[i] & F52_M52;
        const double m = f52_from_int((t * ninv) & F52_M52);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const double h = f52_fma_rz(m, q[j], 0x1p104);
            const double l = f52_fma_rz(m, q[j], (0x1p104 + 0x1p52) - h);
            c[i + j] += f52_bits(l); c[i + j + 1] += f52_bits(h);
        }
        c[i + 1] += (unsigned long long)((long long)c[i] >> 52);
    }
    F52 r;
    unsigned long long carry = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const unsigned long long v = c[5 + k] + carry;
        r.l[k] = f52_from_int(v & F52_M52);
        carry = (unsigned long long)((long long)v >> 52);
    }
    return r;
}

static unsigned long long host_ninv() {
    unsigned long long inv = 1;
    for (int i = 0; i < 7; i++) inv *= 2 - 0x08c16d87cfd47ull * inv;
    return (0 - inv) & F52_M52;
}

__global__ void k_f52_check(const double *in, double *out, int n, unsigned long long ninv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F52 a, b;
    for (int k = 0; k < 5; k++) { a.l[k] = in[(size_t)i * 10 + k]; b.l[k] = in[(size_t)i * 10 + 5 + k]; }
    const F52 r = f52_mul(a, b, ninv);
    for (int k = 0; k < 5; k++) out[(size_t)i * 5 + k] = r.l[k];
}

// role by CTA: (blockIdx.x / sms) % 4 < nfp -> FP64 formulation
__global__ void __launch_bounds__(128) k_mix(uint32_t *out, long long *cyc, int iters_imad, int iters_fp, int sms, int nfp, unsigned long long ninv) {
    const bool fp = (int)((blockIdx.x / sms) & 3) < nfp;
    const long long t0 = clock64();
    uint32_t res = 0;
    if (fp) {
        F52 x, y;
        for (int k = 0; k < 5; k++) { x.l[k] = (double)(0x8000000000000ull + threadIdx.x * 977 + k); y.l[k] = (double)(0x9000000000001ull + blockIdx.x * 31 + k); }
        x.l[4] = 12345.0; y.l[4] = 0x1p44;
        for (int i = 0; i < iters_fp; i++) x = f52_mul(x, y, ninv);
        for (int k = 0; k < 5; k++) res ^= (uint32_t)f52_bits(x.l[k]);
    } else {
        Fq x = Fq::one(), y;
        for (int k = 0; k < 8; k++) y.v[k] = Fq::r2().v[k] ^ (threadIdx.x & 0xff);
        y.v[7] &= 0x0fffffff; x.v[0] += blockIdx.x;
        for (int i = 0; i < iters_imad; i++) x = Fq::mul_impl(x, y);
        for (int k = 0; k < 8; k++) res ^= x.v[k];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 4 + (threadIdx.x >> 5)] = t1 - t0;
}

struct MixResult { double ms, imad_G, fp_G, imad_cyc, fp_cyc; };
static MixResult run_mix(int sms, int ctas_per_sm, int nfp, int iters_imad, int iters_fp, double clock_hz, unsigned long long ninv) {
    const size_t smem = (size_t)(220 * 1024 / ctas_per_sm) & ~(size_t)1023;
    cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int blocks = sms * ctas_per_sm;
    uint32_t *out; long long *cyc; cudaMalloc(&out, (size_t)blocks * 128 * 4); cudaMalloc(&cyc, (size_t)blocks * 4 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<<<blocks, 128, smem>>>(out, cyc, 8, 8, sms, nfp, ninv);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_mix<<<blocks, 128, smem>>>(out, cyc, iters_imad, iters_fp, sms, nfp, ninv);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
    std::vector<long long> h((size_t)blocks * 4);
    cudaMemcpy(h.data(), cyc, h.size() * 8, cudaMemcpyDeviceToHost);
    double ci = 0, cf = 0; long ni = 0, nf = 0;
    for (int b = 0; b < blocks; b++) for (int w = 0; w < 4; w++) { if (((b / sms) & 3) < nfp) { cf += h[b * 4 + w]; nf++; } else { ci += h[b * 4 + w]; ni++; } }
    MixResult r{};
    r.ms = ms; r.imad_cyc = ni ? ci / ni : 0; r.fp_cyc = nf ? cf / nf : 0;
    // per-role rate while the role was running: threads * iterations / (average cycles of its warps / clock)
    r.imad_G = ni ? (double)ni * 32 * iters_imad / (r.imad_cyc / clock_hz) / 1e9 : 0;
    r.fp_G = nf ? (double)nf * 32 * iters_fp / (r.fp_cyc / clock_hz) / 1e9 : 0;
    cudaFree(out); cudaFree(cyc);
    return r;
}

int main(int argc, char **argv) {
    const unsigned long long ninv = host_ninv();
    std::mt19937_64 rng(11);
    auto rnd = [&](F52 &a, unsigned long long ai[5]) { for (int k = 0; k < 5; k++) { ai[k] = rng() & (k == 4 ? ((1ull << 48) - 1) : F52_M52); a.l[k] = (double)ai[k]; } };
    if (argc > 1 && !strcmp(argv[1], "host")) {
        fesetround(FE_TOWARDZERO);
        const int n = argc > 2 ? atoi(argv[2]) : 1000;
        for (int t = 0; t < n; t++) {
            F52 a, b; unsigned long long ai[5], bi[5];
            rnd(a, ai); rnd(b, bi);
            if (t == 0) for (int k = 0; k < 5; k++) { ai[k] = bi[k] = k == 4 ? ((1ull << 48) - 1) : F52_M52; a.l[k] = b.l[k] = (double)ai[k]; }
            if (t == 1) for (int k = 0; k < 5; k++) { ai[k] = 0; a.l[k] = 0; }
            const F52 r = f52_mul(a, b, ninv);
            for (int k = 4; k >= 0; k--) printf("%013llx", ai[k]); printf(" ");
            for (int k = 4; k >= 0; k--) printf("%013llx", bi[k]); printf(" ");
            for (int k = 4; k >= 0; k--) printf("%013llx", (unsigned long long)r.l[k]); printf("\n");
        }
        return 0;
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double clk = p.clockRate * 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    {   // device f52_mul == host f52_mul, bit for bit
        const int n = 1 << 16;
        std::vector<double> in((size_t)n * 10), out((size_t)n * 5), ref((size_t)n * 5);
        fesetround(FE_TOWARDZERO);
        for (int i = 0; i < n; i++) {
            F52 a, b; unsigned long long ai[5], bi[5]; rnd(a, ai); rnd(b, bi);
            for (int k = 0; k < 5; k++) { in[(size_t)i * 10 + k] = a.l[k]; in[(size_t)i * 10 + 5 + k] = b.l[k]; }
            const F52 r = f52_mul(a, b, ninv);
            for (int k = 0; k < 5; k++) ref[(size_t)i * 5 + k] = r.l[k];
        }
        fesetround(FE_TONEAREST);
        double *din, *dout; cudaMalloc(&din, in.size() * 8); cudaMalloc(&dout, out.size() * 8);
        cudaMemcpy(din, in.data(), in.size() * 8, cudaMemcpyHostToDevice);
        k_f52_check<<<n / 128, 128>>>(din, dout, n, ninv);
        cudaMemcpy(out.data(), dout, out.size() * 8, cudaMemcpyDeviceToHost);
        long bad = 0; for (size_t i = 0; i < out.size(); i++) bad += memcmp(&out[i], &ref[i], 8) != 0;
        printf("{\"check\": \"device f52_mul vs host f52_mul\", \"products\": %d, \"mismatching_limbs\": %ld}\n", n, bad);
    }
    for (int ctas : {4, 5, 6, 8}) {
        const int IT = 4096;
        // standalone rates at this residency, then mixes whose iteration counts make both roles end together
        const MixResult a = run_mix(sms, ctas, 0, IT, IT, clk, ninv), f = run_mix(sms, ctas, 4, IT, IT, clk, ninv);
        const double imad_alone = (double)sms * ctas * 128 * IT / (a.ms * 1e-3) / 1e9, fp_alone = (double)sms * ctas * 128 * IT / (f.ms * 1e-3) / 1e9;
        printf("{\"ctas_per_sm\": %d, \"imad_alone_G\": %.2f, \"fp64_alone_G\": %.2f}\n", ctas, imad_alone, fp_alone);
        for (int nfp = 1; nfp <= 3; nfp++) {
            // first guess: each role runs at its standalone per-warp rate; then rebalance once from the measured cycles
            int it_i = IT, it_f = (int)(IT * fp_alone / imad_alone);
            MixResult m = run_mix(sms, ctas, nfp, it_i, it_f, clk, ninv);
            for (int pass = 0; pass < 2; pass++) {
                it_f = (int)(it_f * m.imad_cyc / m.fp_cyc); if (it_f < 16) it_f = 16;
                m = run_mix(sms, ctas, nfp, it_i, it_f, clk, ninv);
            }
            // CTAs per role: roles cycle with period 4 over blockIdx / sms
            int n_f = 0; for (int c = 0; c < ctas; c++) n_f += (c & 3) < nfp;
            const double total = ((double)sms * (ctas - n_f) * 128 * it_i + (double)sms * n_f * 128 * it_f) / (m.ms * 1e-3) / 1e9;
            printf("{\"ctas_per_sm\": %d, \"fp64_ctas_of_4\": %d, \"iters_imad\": %d, \"iters_fp64\": %d, \"ms\": %.3f, \"imad_role_G\": %.2f, \"fp64_role_G\": %.2f, "
                   "\"total_G\": %.2f, \"vs_imad_alone\": %.3f, \"cyc_ratio_fp_over_imad\": %.3f}\n",
                   ctas, nfp, it_i, it_f, m.ms, m.imad_G, m.fp_G, total, total / imad_alone, m.fp_cyc / m.imad_cyc);
            fflush(stdout);
        }
    }
    return 0;
}
