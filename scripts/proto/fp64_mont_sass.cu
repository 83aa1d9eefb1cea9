// Device transcription of fp64_mont.cpp, compiled only to COUNT instructions (nvcc -cubin, cuobjdump -sass): 110 DFMA.RZ + 70 DADD and
// ~260 integer/move instructions per multiplication on sm_100a.  The quotient constant is a placeholder: this file is not a correctness test.
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ long long dbits(double d) { return __double_as_longlong(d); }
__device__ __forceinline__ double from_int52(unsigned long long x) { return __longlong_as_double((long long)(x | 0x4330000000000000ull)) - 0x1p52; }
struct D5 { double l[5]; };
__device__ __forceinline__ void mul_split(double a, double b, long long &hi, long long &lo) {
    const double h = __fma_rz(a, b, 0x1p104);
    const double l = __fma_rz(a, b, (0x1p104 + 0x1p52) - h);
    hi = dbits(h); lo = dbits(l);                       // biases are removed once per column
}
__device__ __noinline__ D5 mont_mul(const D5 a, const D5 b) {
    const unsigned long long M52 = (1ull << 52) - 1;
    const double q[5] = {(double)0x08c16d87cfd47ull, (double)0x916871ca8d3c2ull, (double)0x181585d97816aull, (double)0xa029b85045b68ull, (double)0x30644e72e131ull};
    const unsigned long long Q[5] = {0x08c16d87cfd47ull, 0x916871ca8d3c2ull, 0x181585d97816aull, 0xa029b85045b68ull, 0x30644e72e131ull};
    const double ninv = (double)0x9ffffcd300000001ull;  // placeholder constant of the right magnitude (instruction count only)
    long long col[11];
#pragma unroll
    for (int k = 0; k < 11; k++) col[k] = 0;
    const long long BH = 0x4670000000000000ll, BL = 0x4330000000000000ll;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) { long long hi, lo; mul_split(a.l[i], b.l[j], hi, lo); col[i + j] += lo; col[i + j + 1] += hi; }
    // remove the exponent biases: column k received nlo(k) low halves and nhi(k) high halves
#pragma unroll
    for (int k = 0; k < 10; k++) {
        const int nlo = k < 5 ? k + 1 : 9 - k, nhi = k == 0 ? 0 : (k - 1 < 5 ? k : 10 - k);
        col[k] -= nlo * BL + nhi * BH;
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const unsigned long long t = (unsigned long long)col[i] & M52;
        long long hi, lo; mul_split(from_int52(t), ninv, hi, lo);
        const double m = __longlong_as_double(lo) - 0x1p52;
#pragma unroll
        for (int j = 0; j < 5; j++) { mul_split(m, q[j], hi, lo); col[i + j] += lo - BL; col[i + j + 1] += hi - BH; }
        col[i + 1] += col[i] >> 52;
    }
    unsigned long long r[5]; long long c = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) { const long long v = col[5 + k] + c; r[k] = (unsigned long long)v & M52; c = v >> 52; }
    unsigned long long d[5]; long long br = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) { const long long v = (long long)r[k] - (long long)Q[k] + br; d[k] = (unsigned long long)v & M52; br = v >> 52; }
    D5 out;
#pragma unroll
    for (int k = 0; k < 5; k++) out.l[k] = from_int52(br < 0 ? r[k] : d[k]);
    return out;
}
__global__ void k(double *io, int iters) {
    D5 x, y;
    for (int k2 = 0; k2 < 5; k2++) { x.l[k2] = io[threadIdx.x * 10 + k2]; y.l[k2] = io[threadIdx.x * 10 + 5 + k2]; }
    for (int i = 0; i < iters; i++) x = mont_mul(x, y);
    for (int k2 = 0; k2 < 5; k2++) io[threadIdx.x * 10 + k2] = x.l[k2];
}
