// TEST INFRASTRUCTURE (oracle) -- not product code.
// Shared helpers for the harnesses that wrap the UNMODIFIED reference sources
// (compiled in place from /root/reference by oracle/Makefile).  Nothing here is
// copied from the reference; it only calls its public types.
//
// Wire formats used by every harness entry point (all little-endian):
//   Fr / Fq element : 32 bytes, canonical (NON-Montgomery) value
//   G1 affine       : x(32) y(32); point at infinity = 64 zero bytes
//   G2 affine       : x.c0(32) x.c1(32) y.c0(32) y.c1(32); infinity = 128 zero bytes
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <string>
#include <vector>
#include "libff/algebra/curves/alt_bn128/alt_bn128_pp.hpp"

namespace refio {
typedef libff::alt_bn128_pp ppT;
typedef libff::Fr<ppT> FrT;
typedef libff::Fq<ppT> FqT;
typedef libff::G1<ppT> G1T;
typedef libff::G2<ppT> G2T;

template <typename F> static inline void put_fp(uint8_t *out, const F &v) {
    auto b = v.as_bigint();
    memcpy(out, b.data, 32);
}
template <typename F> static inline F get_fp(const uint8_t *in) {
    libff::bigint<4> b;
    memcpy(b.data, in, 32);
    return F(b);
}
static inline void put_g1(uint8_t *out, const G1T &p) {
    if (p.is_zero()) { memset(out, 0, 64); return; }
    G1T a = p; a.to_affine_coordinates();
    put_fp(out, a.X); put_fp(out + 32, a.Y);
}
static inline G1T get_g1(const uint8_t *in) {
    bool z = true; for (int i = 0; i < 64; i++) if (in[i]) { z = false; break; }
    if (z) return G1T::zero();
    return G1T(get_fp<FqT>(in), get_fp<FqT>(in + 32), FqT::one());
}
static inline void put_g2(uint8_t *out, const G2T &p) {
    if (p.is_zero()) { memset(out, 0, 128); return; }
    G2T a = p; a.to_affine_coordinates();
    put_fp(out, a.X.c0); put_fp(out + 32, a.X.c1); put_fp(out + 64, a.Y.c0); put_fp(out + 96, a.Y.c1);
}
static inline G2T get_g2(const uint8_t *in) {
    bool z = true; for (int i = 0; i < 128; i++) if (in[i]) { z = false; break; }
    if (z) return G2T::zero();
    return G2T(libff::alt_bn128_Fq2(get_fp<FqT>(in), get_fp<FqT>(in + 32)),
               libff::alt_bn128_Fq2(get_fp<FqT>(in + 64), get_fp<FqT>(in + 96)),
               libff::alt_bn128_Fq2::one());
}
static inline bool write_file(const std::string &path, const void *data, size_t len) {
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) return false;
    size_t w = fwrite(data, 1, len, f);
    fclose(f);
    return w == len;
}
static inline bool write_fr_vec(const std::string &path, const std::vector<FrT> &v, size_t count) {
    std::vector<uint8_t> buf(count * 32);
    for (size_t i = 0; i < count; i++) put_fp(&buf[i * 32], v[i]);
    return write_file(path, buf.data(), buf.size());
}
} // namespace refio
