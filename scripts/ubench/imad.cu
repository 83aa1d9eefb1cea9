// Microbenchmarks behind the integer-multiply roofline: issue rate of IMAD, IMAD.WIDE and of the
// whole 254-bit Montgomery multiplication (ff.cuh) as a function of resident warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o imad imad.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../blockmaze_b200/csrc/ff.cuh"
using namespace zk;

__global__ void k_lo(uint32_t *out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 3;
    uint32_t x[8] = {a, a * 3, a ^ b, a + b, a * 5, a * 7, a * 11, a * 13};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(x[k]) : "r"(b));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] ^ x[1] ^ x[2] ^ x[3] ^ x[4] ^ x[5] ^ x[6] ^ x[7];
}
// d(64) = lo(d) * b + d : cannot be strength-reduced, one IMAD.WIDE.U32 per step
__global__ void k_wide(uint32_t *out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 3;
    uint32_t lo[8] = {a, a * 3, a ^ b, a + b, a * 5, a * 7, a * 11, a * 13}, hi[8] = {a, a, a, a, 5, 6, 7, 8};
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int k = 0; k < 8; k++)
                asm volatile("{ .reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t; }"
                             : "+r"(lo[k]), "+r"(hi[k]) : "r"(b));
    }
    uint32_t r = 0;
    for (int k = 0; k < 8; k++) r ^= lo[k] ^ hi[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// carry-chained wide: mad.lo.cc / madc.hi.cc pairs as in ff.cuh (IMAD.WIDE.U32.X with predicate carry in/out)
__global__ void k_widex(uint32_t *out, int iters) {
    uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 3;
    uint32_t v[18];
    for (int k = 0; k < 18; k++) v[k] = a * (k + 1);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(v[0]) : "r"(v[16]), "r"(b));
            asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(v[1]) : "r"(v[16]), "r"(b));
#pragma unroll
            for (int k = 2; k < 16; k += 2) {
                asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(v[k]) : "r"(v[17]), "r"(b));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(v[k + 1]) : "r"(v[17]), "r"(b));
            }
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(v[16]));
            v[17] ^= v[3];
        }
    }
    uint32_t r = 0;
    for (int k = 0; k < 18; k++) r ^= v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// FP64 FMA issue rate (data point for a 52-bit-limb floating-point multiplication, DESIGN.md "what comes next")
__global__ void k_dfma(uint32_t *out, int iters) {
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - blockIdx.x * 1e-9;
    double x[8];
    for (int k = 0; k < 8; k++) x[k] = a * (k + 1);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = __fma_rz(x[k], b, a);
    }
    double r = 0;
    for (int k = 0; k < 8; k++) r += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double2ll_rz(r);
}
// 64-bit integer add chains (IADD3 + IADD3.X pairs on the ALU pipe)
__global__ void k_add64(uint32_t *out, int iters) {
    unsigned long long x[8], b = blockIdx.x * 40503ull + threadIdx.x + 3;
    for (int k = 0; k < 8; k++) x[k] = b * (k + 1);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("add.u64 %0, %0, %1;" : "+l"(x[k]) : "l"(x[(k + 1) & 7]));
    }
    unsigned long long r = 0;
    for (int k = 0; k < 8; k++) r ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(r ^ (r >> 32));
}
template <int ILP> __global__ void k_modmul(uint32_t *out, int iters) {
    Fq x[ILP], y;
    for (int k = 0; k < 8; k++) { y.v[k] = Fq::r2().v[k] ^ (threadIdx.x & 0xff); }
    y.v[7] &= 0x0fffffff;
    for (int j = 0; j < ILP; j++) { x[j] = Fq::one(); x[j].v[0] += blockIdx.x + j; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = Fq::mul_impl(x[j], y);
    }
    uint32_t r = 0;
    for (int j = 0; j < ILP; j++) for (int k = 0; k < 8; k++) r ^= x[j].v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_modmul_call(uint32_t *out, int iters) {
    Fq x = Fq::one(), y;
    for (int k = 0; k < 8; k++) { y.v[k] = Fq::r2().v[k] ^ (threadIdx.x & 0xff); }
    y.v[7] &= 0x0fffffff;
    x.v[0] += blockIdx.x;
    for (int i = 0; i < iters; i++) x = x * y;
    uint32_t r = 0;
    for (int k = 0; k < 8; k++) r ^= x.v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K> static double run(K kern, int warps_per_sm, int iters, double ops_per_thread_iter, int sms) {
    // one CTA of 128 threads = 4 warps; dynamic shared memory caps the CTAs per SM
    const int ctas_per_sm = warps_per_sm / 4;
    const size_t smem = (size_t)(220 * 1024 / ctas_per_sm) & ~(size_t)1023;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int blocks = sms * ctas_per_sm;
    uint32_t *out; cudaMalloc(&out, (size_t)blocks * 128 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<<<blocks, 128, smem>>>(out, 8);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<blocks, 128, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
    cudaFree(out);
    return (double)blocks * 128 * iters * ops_per_thread_iter / (ms * 1e-3);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    const int occ[] = {4, 8, 16, 24, 32, 48, 64};
    for (int w : occ) {
        printf("{\"warps_per_sm\": %d, \"imad_lo_T\": %.3f, \"imad_wide_T\": %.3f, \"imad_widex_T\": %.3f, "
               "\"modmul_inl_G\": %.2f, \"modmul_inl_ilp2_G\": %.2f, \"modmul_call_G\": %.2f, \"dfma_T\": %.3f, \"add64_T\": %.3f}\n", w,
               run(k_lo, w, 2048, 64, sms) / 1e12, run(k_wide, w, 2048, 64, sms) / 1e12, run(k_widex, w, 2048, 32, sms) / 1e12,
               run(k_modmul<1>, w, 2048, 1, sms) / 1e9, run(k_modmul<2>, w, 1024, 2, sms) / 1e9, run(k_modmul_call, w, 2048, 1, sms) / 1e9,
               run(k_dfma, w, 2048, 64, sms) / 1e12, run(k_add64, w, 2048, 64, sms) / 1e12);
        fflush(stdout);
    }
    return 0;
}
