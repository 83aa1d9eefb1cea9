// The 24 792 internal variables of one sha256_compression_function_gadget as a function of its 64-round trace
// (libsnark/gadgetlib1/gadgets/hashes/sha256/sha256_gadget.tcc:20-140, sha256_components.tcc, sha256_aux.tcc; allocation order of the
// gadget's constructor).  Shared by the GPU kernel that expands the runs (prover.cu sha256_witness_kernel) and by the host (witness.hpp
// Assignment::expand, the checker of that kernel), so both read the layout from one place:
//   [0, 64)                     packed_W[i]
//   64 + 152*(i-16), i = 16..63 message schedule: sigma0, sigma1, 32 bits sigma0, 29 bits of rotr7^rotr18, 32 bits sigma1, 22 bits of
//                               rotr17^rotr19, the unreduced sum W[i-16]+sigma0+W[i-7]+sigma1 and its 34 bits
//   7360 + 272*i, i = 0..63     round: 32 bits of the new_a sum, 32 bits of the new_e sum, Sigma0, Sigma1, bits Sigma0, bits of its first XOR,
//                               bits Sigma1, bits of its first XOR, ch + 32 bits, maj + 32 bits, d, h, the two unreduced sums, new_a, new_e,
//                               3 + 3 overflow bits
//   24768 + o, o = 0..7         unreduced output word o (chaining value + last registers); + 8: reduced; + 16: its carry bit
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define ZKW_HD __host__ __device__ inline
#else
#define ZKW_HD inline
#endif

namespace zkw {

constexpr uint32_t SHA_RUN_VARS = 24792;
struct ShaTrace { uint64_t W[64], sch[48][5], rnd[64][10], fin[8]; };

ZKW_HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
ZKW_HD uint32_t sha_k(int i) {
    const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    return K[i];
}
// the trace of one compression: w = the 16 message words, h = the incoming chaining value
ZKW_HD void sha_trace_build(const uint32_t w[16], const uint32_t h8[8], ShaTrace &T) {
    for (int i = 0; i < 16; i++) T.W[i] = w[i];
    for (int i = 16; i < 64; i++) {
        const uint32_t x0 = (uint32_t)T.W[i - 15], x1 = (uint32_t)T.W[i - 2];
        const uint32_t ab0 = sha_rotr(x0, 7) ^ sha_rotr(x0, 18), ab1 = sha_rotr(x1, 17) ^ sha_rotr(x1, 19);
        const uint64_t sig0 = ab0 ^ (x0 >> 3), sig1 = ab1 ^ (x1 >> 10);
        const uint64_t unred = sig0 + sig1 + T.W[i - 16] + T.W[i - 7];
        T.sch[i - 16][0] = sig0; T.sch[i - 16][1] = ab0; T.sch[i - 16][2] = sig1; T.sch[i - 16][3] = ab1; T.sch[i - 16][4] = unred;
        T.W[i] = unred & 0xffffffffull;
    }
    uint32_t a = h8[0], b = h8[1], c = h8[2], d = h8[3], e = h8[4], f = h8[5], g = h8[6], h = h8[7];
    for (int i = 0; i < 64; i++) {
        const uint32_t t0 = sha_rotr(a, 2) ^ sha_rotr(a, 13), t1 = sha_rotr(e, 6) ^ sha_rotr(e, 11);
        const uint64_t S0 = t0 ^ sha_rotr(a, 22), S1 = t1 ^ sha_rotr(e, 25);
        const uint64_t ch = (e & f) | (~e & g), mj = (a & b) | (a & c) | (b & c);
        const uint64_t ua = (uint64_t)h + S1 + ch + sha_k(i) + T.W[i] + S0 + mj;
        const uint64_t ue = (uint64_t)d + h + S1 + ch + sha_k(i) + T.W[i];
        uint64_t *q = T.rnd[i];
        q[0] = S0; q[1] = t0; q[2] = S1; q[3] = t1; q[4] = ch; q[5] = mj; q[6] = d; q[7] = h; q[8] = ua; q[9] = ue;
        h = g; g = f; f = e; e = (uint32_t)ue; d = c; c = b; b = a; a = (uint32_t)ua;
    }
    const uint32_t last[8] = {a, b, c, d, e, f, g, h};
    for (int o = 0; o < 8; o++) T.fin[o] = (uint64_t)h8[o] + last[o];
}
// value of variable `v` (0 <= v < SHA_RUN_VARS) of the run
ZKW_HD uint64_t sha_trace_value(const ShaTrace &T, uint32_t v) {
    if (v < 64) return T.W[v];
    if (v < 7360) {
        const uint32_t i = (v - 64) / 152, o = (v - 64) % 152;
        const uint64_t *q = T.sch[i];
        if (o == 0) return q[0];
        if (o == 1) return q[2];
        if (o < 34) return (q[0] >> (o - 2)) & 1;
        if (o < 63) return (q[1] >> (o - 34)) & 1;
        if (o < 95) return (q[2] >> (o - 63)) & 1;
        if (o < 117) return (q[3] >> (o - 95)) & 1;
        if (o == 117) return q[4];
        return (q[4] >> (o - 118)) & 1;
    }
    if (v < 24768) {
        const uint32_t i = (v - 7360) / 272, o = (v - 7360) % 272;
        const uint64_t *q = T.rnd[i];
        if (o < 32) return (q[8] >> o) & 1;
        if (o < 64) return (q[9] >> (o - 32)) & 1;
        if (o == 64) return q[0];
        if (o == 65) return q[2];
        if (o < 98) return (q[0] >> (o - 66)) & 1;
        if (o < 130) return (q[1] >> (o - 98)) & 1;
        if (o < 162) return (q[2] >> (o - 130)) & 1;
        if (o < 194) return (q[3] >> (o - 162)) & 1;
        if (o == 194) return q[4];
        if (o < 227) return (q[4] >> (o - 195)) & 1;
        if (o == 227) return q[5];
        if (o < 260) return (q[5] >> (o - 228)) & 1;
        if (o == 260) return q[6];
        if (o == 261) return q[7];
        if (o == 262) return q[8];
        if (o == 263) return q[9];
        if (o == 264) return q[8] & 0xffffffffull;
        if (o == 265) return q[9] & 0xffffffffull;
        if (o < 269) return (q[8] >> (32 + o - 266)) & 1;
        return (q[9] >> (32 + o - 269)) & 1;
    }
    const uint32_t o = v - 24768;
    if (o < 8) return T.fin[o];
    if (o < 16) return T.fin[o - 8] & 0xffffffffull;
    return (T.fin[o - 16] >> 32) & 1;
}

} // namespace zkw
