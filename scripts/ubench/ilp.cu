// Latency of dependent vs independent 254-bit Montgomery multiplications in ONE warp (the regime of the MSM tails: chains of
// dependent point additions with nothing else to hide latency).  Prints clocks per multiplication for
//   serial  : x = x*y, four times, each depending on the last          (out-of-line call, as the tail kernels do today)
//   ilp2/4  : 2 / 4 independent products written back to back, inlined  (does ptxas interleave the carry chains?)
//   add     : one XYZZ += XYZZ with out-of-line multiplications (ec.cuh add) vs the grouped inline version below
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I blockmaze_b200/csrc scripts/ubench/ilp.cu -o scripts/ubench/ilp
#include <cstdio>
#include <cuda_runtime.h>
#include "ec.cuh"
using namespace zk;

// ---- the quad-cooperative point operations that were measured (and dropped, see profiles/r02_notes.md): kept here with the benchmark -----
template <class F> __device__ __forceinline__ F quad_bcast(const F &v, int src) {
    F r; const uint32_t *s = reinterpret_cast<const uint32_t *>(&v); uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = __shfl_sync(0xffffffffu, s[i], src, 4);
    return r;
}
template <class F> __device__ __forceinline__ F quad_sel(int l, const F &a, const F &b, const F &c, const F &d) {
    F r; const uint32_t *pa = reinterpret_cast<const uint32_t *>(&a), *pb = reinterpret_cast<const uint32_t *>(&b);
    const uint32_t *pc = reinterpret_cast<const uint32_t *>(&c), *pd = reinterpret_cast<const uint32_t *>(&d); uint32_t *o = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(F) / 4); i++) o[i] = l == 0 ? pa[i] : l == 1 ? pb[i] : l == 2 ? pc[i] : pd[i];
    return r;
}
template <class P> __device__ __forceinline__ P point_sel(bool take_b, const P &a, const P &b) {
    P r; const uint32_t *pa = reinterpret_cast<const uint32_t *>(&a), *pb = reinterpret_cast<const uint32_t *>(&b); uint32_t *o = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(P) / 4); i++) o[i] = take_b ? pb[i] : pa[i];
    return r;
}
template <class F> __device__ __forceinline__ void quad_add(XYZZ<F> &a, const XYZZ<F> &o) {      // all 32 lanes call it together
    const bool o_inf = o.is_inf(), a_inf = a.is_inf();
    const int l = threadIdx.x & 3;
    F m = quad_sel(l, a.X, o.X, a.Y, o.Y) * quad_sel(l, o.ZZ, a.ZZ, o.ZZZ, a.ZZZ);            // U1 | U2 | S1 | S2
    const F U1 = quad_bcast(m, 0), S1 = quad_bcast(m, 2);
    const F Pp = quad_bcast(m, 1) - U1, R = quad_bcast(m, 3) - S1;
    m = quad_sel(l, Pp, R, a.ZZ, a.ZZZ) * quad_sel(l, Pp, R, o.ZZ, o.ZZZ);                    // PP | RR | ZZ1*ZZ2 | ZZZ1*ZZZ2
    const F PP = quad_bcast(m, 0), RR = quad_bcast(m, 1);
    const F m3 = quad_sel(l, Pp, U1, m, Pp) * PP;                                             // PPP | Q | ZZ3 | PPP
    const F PPP = quad_bcast(m3, 0), Q = quad_bcast(m3, 1);
    XYZZ<F> r;
    r.X = RR - PPP - Q.dbl();
    const F m4 = quad_sel(l, R, S1, m, m) * quad_sel(l, Q - r.X, PPP, PPP, PPP);              // R(Q-X3) | S1*PPP | - | ZZZ3
    r.Y = quad_bcast(m4, 0) - quad_bcast(m4, 1);
    r.ZZ = quad_bcast(m3, 2);
    r.ZZZ = quad_bcast(m4, 3);
    const bool degenerate = !o_inf && !a_inf && Pp.is_zero();
    if (__any_sync(0xffffffffu, degenerate)) { if (degenerate) { if (R.is_zero()) r = a.dbl(); else r = XYZZ<F>::inf(); } __syncwarp(); }
    a = point_sel(a_inf, point_sel(o_inf, r, a), o);
}
template <class F> __device__ __forceinline__ void quad_dbl(XYZZ<F> &a) {
    const int l = threadIdx.x & 3;
    const F U = a.Y.dbl();
    F m = quad_sel(l, U, a.X, U, U);
    m = m * m;                                                                                // V | XX | - | -
    const F V = quad_bcast(m, 0), XX = quad_bcast(m, 1);
    const F M = XX.dbl() + XX;
    const F m2 = quad_sel(l, U, a.X, V, M) * quad_sel(l, V, V, a.ZZ, M);                      // W | S | ZZ3 | M^2
    const F W = quad_bcast(m2, 0), S = quad_bcast(m2, 1);
    const F X3 = quad_bcast(m2, 3) - S.dbl();
    const F m3 = quad_sel(l, M, W, W, W) * quad_sel(l, S - X3, a.Y, a.ZZZ, a.ZZZ);            // M(S-X3) | W*Y | ZZZ3 | -
    a.X = X3; a.Y = quad_bcast(m3, 0) - quad_bcast(m3, 1); a.ZZ = quad_bcast(m2, 2); a.ZZZ = quad_bcast(m3, 2);
}

__global__ void k_serial(Fq *io, int iters, long long *clk) {
    Fq x = io[threadIdx.x], y = io[32 + threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { x = x * y; x = x * y; x = x * y; x = x * y; }
    long long t1 = clock64();
    io[threadIdx.x] = x; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_serial_inl(Fq *io, int iters, long long *clk) {
    Fq x = io[threadIdx.x], y = io[32 + threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { x = Fq::mul_impl(x, y); x = Fq::mul_impl(x, y); x = Fq::mul_impl(x, y); x = Fq::mul_impl(x, y); }
    long long t1 = clock64();
    io[threadIdx.x] = x; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_ilp2(Fq *io, int iters, long long *clk) {
    Fq a = io[threadIdx.x], b = io[32 + threadIdx.x], y = io[64 + threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { a = Fq::mul_impl(a, y); b = Fq::mul_impl(b, y); a = Fq::mul_impl(a, y); b = Fq::mul_impl(b, y); }
    long long t1 = clock64();
    io[threadIdx.x] = a + b; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_ilp4(Fq *io, int iters, long long *clk) {
    Fq a = io[threadIdx.x], b = io[32 + threadIdx.x], c = io[64 + threadIdx.x], d = io[96 + threadIdx.x], y = io[128 + threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { a = Fq::mul_impl(a, y); b = Fq::mul_impl(b, y); c = Fq::mul_impl(c, y); d = Fq::mul_impl(d, y); }
    long long t1 = clock64();
    io[threadIdx.x] = a + b + c + d; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_add_call(XYZZ<Fq> *io, int iters, long long *clk) {
    XYZZ<Fq> a = io[threadIdx.x], o = io[32 + threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) a.add(o);
    long long t1 = clock64();
    io[threadIdx.x] = a; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
// the same addition with the multiplications inlined and grouped by dependency level
__device__ __noinline__ void add_grouped(XYZZ<Fq> &a, const XYZZ<Fq> &o) {
    auto M = [](const Fq &x, const Fq &y) { return Fq::mul_impl(x, y); };
    const Fq U1 = M(a.X, o.ZZ), U2 = M(o.X, a.ZZ), S1 = M(a.Y, o.ZZZ), S2 = M(o.Y, a.ZZZ);
    const Fq P = U2 - U1, R = S2 - S1;
    const Fq PP = M(P, P), RR = M(R, R), Z2 = M(a.ZZ, o.ZZ), Z3 = M(a.ZZZ, o.ZZZ);
    const Fq PPP = M(P, PP), Q = M(U1, PP), ZZ3 = M(Z2, PP);
    const Fq X3 = RR - PPP - Q.dbl();
    const Fq t1 = M(R, Q - X3), t2 = M(S1, PPP), ZZZ3 = M(Z3, PPP);
    a.X = X3; a.Y = t1 - t2; a.ZZ = ZZ3; a.ZZZ = ZZZ3;
}
__global__ void k_add_grouped(XYZZ<Fq> *io, int iters, long long *clk) {
    XYZZ<Fq> a = io[threadIdx.x], o = io[32 + threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) add_grouped(a, o);
    long long t1 = clock64();
    io[threadIdx.x] = a; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_quad_add(XYZZ<Fq> *io, int iters, long long *clk) {
    XYZZ<Fq> a = io[threadIdx.x >> 2], o = io[32 + (threadIdx.x >> 2)];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) quad_add(a, o);
    long long t1 = clock64();
    io[threadIdx.x] = a; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_quad_dbl(XYZZ<Fq> *io, int iters, long long *clk) {
    XYZZ<Fq> a = io[threadIdx.x >> 2];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) quad_dbl(a);
    long long t1 = clock64();
    io[threadIdx.x] = a; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_dbl_call(XYZZ<Fq> *io, int iters, long long *clk) {
    XYZZ<Fq> a = io[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) a = a.dbl();
    long long t1 = clock64();
    io[threadIdx.x] = a; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
__global__ void k_shfl(uint32_t *io, int iters, long long *clk) {
    uint32_t v[8]; for (int k = 0; k < 8; k++) v[k] = io[threadIdx.x * 8 + k];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { for (int k = 0; k < 8; k++) v[k] = __shfl_sync(0xffffffffu, v[k], (i + k) & 3, 4) + 1; }
    long long t1 = clock64();
    for (int k = 0; k < 8; k++) io[threadIdx.x * 8 + k] = v[k]; if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
    Fq *io; long long *clk; cudaMalloc(&io, 4096 * 32); cudaMalloc(&clk, 8); cudaMemset(io, 0x11, 4096 * 32);
    const int iters = 2000; long long h;
    for (int rep = 0; rep < 2; rep++) {
#define RUN(name, kern, arg, per) kern<<<1, 32>>>(arg, iters, clk); cudaDeviceSynchronize(); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); \
        if (rep) printf("{\"bench\": \"%s\", \"clk_per_unit\": %.1f, \"err\": \"%s\"}\n", name, (double)h / iters / per, cudaGetErrorString(cudaGetLastError()));
        RUN("serial_call_per_mul", k_serial, io, 4)
        RUN("serial_inline_per_mul", k_serial_inl, io, 4)
        RUN("ilp2_per_mul", k_ilp2, io, 4)
        RUN("ilp4_per_mul", k_ilp4, io, 4)
        RUN("xyzz_add_call_per_add", k_add_call, (XYZZ<Fq> *)io, 1)
        RUN("xyzz_add_grouped_per_add", k_add_grouped, (XYZZ<Fq> *)io, 1)
        RUN("xyzz_dbl_call_per_dbl", k_dbl_call, (XYZZ<Fq> *)io, 1)
        RUN("quad_add_per_add", k_quad_add, (XYZZ<Fq> *)io, 1)
        RUN("quad_dbl_per_dbl", k_quad_dbl, (XYZZ<Fq> *)io, 1)
        RUN("bcast_8_words", k_shfl, (uint32_t *)io, 1)
    }
    return 0;
}
