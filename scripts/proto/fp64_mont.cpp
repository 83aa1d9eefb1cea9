// Prototype (host, not part of the library): BN254 base-field Montgomery multiplication on 5 x 52-bit limbs held in doubles, the
// formulation DESIGN.md "what comes next" proposes for the FP64 pipe of the B200.  Every limb product is split EXACTLY into its high
// and low 52 bits with two fused multiply-adds in round-toward-zero mode (CUDA: __fma_rz) and one subtraction; the halves are summed
// as 64-bit integers taken from the doubles' bit patterns.  R = 2^260.
//   g++ -O2 -std=c++17 -frounding-math -o fp64_mont fp64_mont.cpp && ./fp64_mont 20000 | python3 fp64_mont_check.py
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

static const double C1 = 0x1p104, C2 = 0x1p104 + 0x1p52, TWO52 = 0x1p52;
static const uint64_t M52 = (1ull << 52) - 1;
static inline int64_t bits(double d) { int64_t r; memcpy(&r, &d, 8); return r; }
static inline double from_int52(uint64_t x) { uint64_t u = x | 0x4330000000000000ull; double d; memcpy(&d, &u, 8); return d - TWO52; }

// exact split of a*b (a, b integers < 2^52 held in doubles): hi = floor(ab / 2^52), lo = ab mod 2^52
static inline void mul_split(double a, double b, int64_t &hi, int64_t &lo) {
    const double h = std::fma(a, b, C1);            // = 2^104 + floor(ab / 2^52) * 2^52 under round-toward-zero
    const double l = std::fma(a, b, C2 - h);        // = 2^52 + (ab mod 2^52), exact
    hi = bits(h) - bits(C1);
    lo = bits(l) - bits(TWO52);
}

// q = 21888242871839275222246405745257275088696311157297823662689037894645226208583 in 52-bit limbs
static const uint64_t Q52[5] = {0x08c16d87cfd47ull, 0x916871ca8d3c2ull, 0x181585d97816aull, 0xa029b85045b68ull, 0x30644e72e131ull};
static uint64_t NINV52;                             // -q^-1 mod 2^52

struct D5 { double l[5]; };

static D5 mont_mul(const D5 &a, const D5 &b) {
    int64_t col[11] = {0};
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) { int64_t hi, lo; mul_split(a.l[i], b.l[j], hi, lo); col[i + j] += lo; col[i + j + 1] += hi; }
    double qd[5];
    for (int j = 0; j < 5; j++) qd[j] = from_int52(Q52[j]);
    const double ninv = from_int52(NINV52);
    for (int i = 0; i < 5; i++) {
        const uint64_t t = (uint64_t)col[i] & M52;
        int64_t hi, lo; mul_split(from_int52(t), ninv, hi, lo);
        const double m = from_int52((uint64_t)lo);                      // m = t * (-q^-1) mod 2^52
        for (int j = 0; j < 5; j++) { mul_split(m, qd[j], hi, lo); col[i + j] += lo; col[i + j + 1] += hi; }
        col[i + 1] += col[i] >> 52;                                     // column i is now a multiple of 2^52
    }
    uint64_t r[5]; int64_t c = 0;
    for (int k = 0; k < 5; k++) { const int64_t v = col[5 + k] + c; r[k] = (uint64_t)v & M52; c = v >> 52; }
    // r < 2q: subtract q once if needed
    uint64_t d[5]; int64_t br = 0;
    for (int k = 0; k < 5; k++) { const int64_t v = (int64_t)r[k] - (int64_t)Q52[k] + br; d[k] = (uint64_t)v & M52; br = v >> 52; }
    D5 out;
    for (int k = 0; k < 5; k++) out.l[k] = from_int52(br < 0 ? r[k] : d[k]);
    return out;
}

int main(int argc, char **argv) {
    fesetround(FE_TOWARDZERO);
    uint64_t inv = 1;                                                  // Newton: inv = q^-1 mod 2^64, then negate and mask
    for (int i = 0; i < 7; i++) inv *= 2 - Q52[0] * inv;               // only the low limb of q matters mod 2^52
    NINV52 = (0 - inv) & M52;
    const int n = argc > 1 ? atoi(argv[1]) : 1000;
    std::mt19937_64 rng(7);
    for (int t = 0; t < n; t++) {
        D5 a, b; uint64_t ai[5], bi[5];
        for (int k = 0; k < 5; k++) { ai[k] = rng() & (k == 4 ? 0x1fffffffffffull : M52); bi[k] = rng() & (k == 4 ? 0x1fffffffffffull : M52); }
        if (t == 0) { for (int k = 0; k < 5; k++) { ai[k] = Q52[k]; bi[k] = Q52[k]; } ai[0] -= 1; bi[0] -= 1; }        // (q-1)^2
        if (t == 1) { for (int k = 0; k < 5; k++) { ai[k] = 0; bi[k] = M52; } bi[4] = 0x1fffffffffffull; }
        for (int k = 0; k < 5; k++) { a.l[k] = from_int52(ai[k]); b.l[k] = from_int52(bi[k]); }
        const D5 r = mont_mul(a, b);
        for (int k = 4; k >= 0; k--) printf("%013llx", (unsigned long long)ai[k]); printf(" ");
        for (int k = 4; k >= 0; k--) printf("%013llx", (unsigned long long)bi[k]); printf(" ");
        for (int k = 4; k >= 0; k--) printf("%013llx", (unsigned long long)(uint64_t)r.l[k]); printf("\n");
    }
    return 0;
}
