// Parser for BlockMaze's proving-key files (/usr/local/prfKey/<circuit>pk.txt).
//
// Grammar (reference writers: r1cs_gg_ppzksnark.tcc:52-66 proving key, alt_bn128_g1.cpp:404-418,467-476 and
// alt_bn128_g2.cpp:418-431 points [binary, Montgomery, compressed], sparse_vector.tcc:272-288, knowledge_commitment.tcc:121-125,
// r1cs.tcc:242-254 + variable.tcc:411-421 constraint system [ASCII decimal]):
//   G1 := zero('0'|'1') x[32 LE, Montgomery] ylsb('0'|'1')          34 bytes
//   G2 := zero x.c0[32] x.c1[32] ylsb(of canonical y.c0)             66 bytes
//   pk := G1 \n G1 \n G2 \n G1 \n G2 \n                              alpha_g1 beta_g1 beta_g2 delta_g1 delta_g2
//         dec \n G1{n+1}                                             A_query
//         dec \n dec(k) \n (dec \n){k} dec(k) \n (G2 ' ' G1 \n){k}   B_query (sparse)
//         dec \n G1{m-1}                                             H_query
//         dec \n G1{n-inputs}                                        L_query
//         dec(inputs) \n dec(aux) \n dec(#constraints) \n            constraint system
//         per constraint, a then b then c:  dec(#terms) \n (dec(index) \n dec(coeff) \n){#terms}
// The parser keeps points COMPRESSED (x + flags); y is recovered on the GPU (one Fq / Fq2 square root per point,
// pk_decompress kernels) instead of the reference's 40-110 s host loop (alt_bn128_g1.cpp:420-465).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include "host_field.hpp"

namespace zkpk {

struct CompressedG1 { uint8_t x[32]; uint8_t flags; };   // flags bit0 = y lsb, bit1 = infinity
struct CompressedG2 { uint8_t x[64]; uint8_t flags; };

struct Csr {
    std::vector<uint32_t> rowptr, col, coef;   // coef = index into ParsedPk::coef_dict
};

struct ParsedPk {
    CompressedG1 alpha_g1, beta_g1, delta_g1;
    CompressedG2 beta_g2, delta_g2;
    std::vector<CompressedG1> A, B_g1, H, L;
    std::vector<CompressedG2> B_g2;
    std::vector<uint32_t> B_idx;
    uint64_t B_domain = 0;
    uint64_t num_inputs = 0, num_aux = 0, num_constraints = 0;
    Csr a, b, c;
    std::vector<zkh::HFr> coef_dict;           // Montgomery; [0] = 1, [1] = -1
    std::string error;
};

class Cursor {
  public:
    Cursor(const uint8_t *d, size_t n) : d_(d), n_(n), p_(0) {}
    bool ok() const { return !fail_; }
    size_t pos() const { return p_; }
    bool expect(char ch) { if (p_ < n_ && d_[p_] == (uint8_t)ch) { p_++; return true; } fail_ = true; return false; }
    // ASCII decimal terminated by ' ' or '\n' (consumed); value must fit in 64 bits
    uint64_t dec() {
        uint64_t v = 0; size_t s = p_;
        while (p_ < n_ && d_[p_] >= '0' && d_[p_] <= '9') { v = v * 10 + (d_[p_] - '0'); p_++; }
        if (p_ == s || p_ >= n_ || (d_[p_] != '\n' && d_[p_] != ' ')) { fail_ = true; return 0; }
        p_++;
        return v;
    }
    // decimal token as text (for field elements)
    bool dec_token(const char *&s, size_t &len) {
        size_t b = p_;
        while (p_ < n_ && d_[p_] >= '0' && d_[p_] <= '9') p_++;
        if (p_ == b || p_ >= n_ || (d_[p_] != '\n' && d_[p_] != ' ')) { fail_ = true; return false; }
        s = (const char *)d_ + b; len = p_ - b; p_++;
        return true;
    }
    bool g1(CompressedG1 &o) {
        if (p_ + 34 > n_) { fail_ = true; return false; }
        uint8_t z = d_[p_], l = d_[p_ + 33];
        if ((z != '0' && z != '1') || (l != '0' && l != '1')) { fail_ = true; return false; }
        memcpy(o.x, d_ + p_ + 1, 32);
        o.flags = (uint8_t)((l - '0') | ((z - '0') << 1));
        p_ += 34;
        return true;
    }
    bool g2(CompressedG2 &o) {
        if (p_ + 66 > n_) { fail_ = true; return false; }
        uint8_t z = d_[p_], l = d_[p_ + 65];
        if ((z != '0' && z != '1') || (l != '0' && l != '1')) { fail_ = true; return false; }
        memcpy(o.x, d_ + p_ + 1, 64);
        o.flags = (uint8_t)((l - '0') | ((z - '0') << 1));
        p_ += 66;
        return true;
    }
  private:
    const uint8_t *d_; size_t n_, p_; bool fail_ = false;
};

inline bool parse_g1_vec(Cursor &c, std::vector<CompressedG1> &v) {
    uint64_t n = c.dec();
    if (!c.ok() || n > (1ull << 28)) return false;
    v.resize(n);
    for (uint64_t i = 0; i < n; i++) if (!c.g1(v[i])) return false;
    return true;
}

inline bool parse_pk(const uint8_t *data, size_t len, ParsedPk &pk) {
    Cursor c(data, len);
#define ZK_REQ(x) do { if (!(x) || !c.ok()) { pk.error = "pk parse error near byte " + std::to_string(c.pos()); return false; } } while (0)
    ZK_REQ(c.g1(pk.alpha_g1)); ZK_REQ(c.expect('\n'));
    ZK_REQ(c.g1(pk.beta_g1)); ZK_REQ(c.expect('\n'));
    ZK_REQ(c.g2(pk.beta_g2)); ZK_REQ(c.expect('\n'));
    ZK_REQ(c.g1(pk.delta_g1)); ZK_REQ(c.expect('\n'));
    ZK_REQ(c.g2(pk.delta_g2)); ZK_REQ(c.expect('\n'));
    ZK_REQ(parse_g1_vec(c, pk.A));
    pk.B_domain = c.dec();
    uint64_t k = c.dec();
    ZK_REQ(k <= (1ull << 28));
    pk.B_idx.resize(k);
    for (uint64_t i = 0; i < k; i++) { const uint64_t idx = c.dec(); ZK_REQ(idx < pk.B_domain && pk.B_domain <= (1ull << 28)); pk.B_idx[i] = (uint32_t)idx; }
    uint64_t k2 = c.dec();
    ZK_REQ(k2 == k);
    pk.B_g1.resize(k); pk.B_g2.resize(k);
    for (uint64_t i = 0; i < k; i++) {
        ZK_REQ(c.g2(pk.B_g2[i])); ZK_REQ(c.expect(' '));
        ZK_REQ(c.g1(pk.B_g1[i])); ZK_REQ(c.expect('\n'));
    }
    ZK_REQ(parse_g1_vec(c, pk.H));
    ZK_REQ(parse_g1_vec(c, pk.L));
    pk.num_inputs = c.dec(); pk.num_aux = c.dec(); pk.num_constraints = c.dec();
    ZK_REQ(pk.num_constraints <= (1ull << 28));

    // coefficient dictionary keyed by the decimal text, so each distinct value is converted once
    std::unordered_map<std::string, uint32_t> dict;
    pk.coef_dict.clear();
    pk.coef_dict.push_back(zkh::HFr::one());
    pk.coef_dict.push_back(zkh::HFr::one().neg());
    dict.emplace("1", 0u);
    dict.emplace("21888242871839275222246405745257275088548364400416034343698204186575808495616", 1u);
    Csr *mats[3] = {&pk.a, &pk.b, &pk.c};
    for (auto *m : mats) { m->rowptr.assign(1, 0); m->rowptr.reserve(pk.num_constraints + 1); }
    const uint64_t nvars = pk.num_inputs + pk.num_aux;
    std::string key;
    for (uint64_t i = 0; i < pk.num_constraints; i++) {
        for (int w = 0; w < 3; w++) {
            uint64_t nt = c.dec();
            ZK_REQ(nt <= nvars + 1);
            for (uint64_t t = 0; t < nt; t++) {
                uint64_t idx = c.dec();
                const char *s; size_t sl;
                ZK_REQ(c.dec_token(s, sl));
                ZK_REQ(idx <= nvars);
                uint32_t ci;
                if (sl == 1 && s[0] == '1') ci = 0;
                else {
                    key.assign(s, sl);
                    auto it = dict.find(key);
                    if (it == dict.end()) {
                        zkh::HFr v;
                        ZK_REQ(zkh::HFr::from_dec(s, sl, v));
                        ci = (uint32_t)pk.coef_dict.size();
                        pk.coef_dict.push_back(v);
                        dict.emplace(key, ci);
                    } else ci = it->second;
                }
                mats[w]->col.push_back((uint32_t)idx);
                mats[w]->coef.push_back(ci);
            }
            mats[w]->rowptr.push_back((uint32_t)mats[w]->col.size());
        }
    }
#undef ZK_REQ
    return true;
}

} // namespace zkpk
