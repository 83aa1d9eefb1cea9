// Key generator: r1cs_gg_ppzksnark_generator (libsnark/zk_proof_systems/ppzksnark/r1cs_gg_ppzksnark/r1cs_gg_ppzksnark.tcc:211-388)
// for a constraint system taken from an existing proving-key file, writing <circuit>pk.txt / <circuit>vk.txt in the reference's
// on-disk format.  SURVEY.md section 8(f) rank 4: 35-106 s per circuit on the CPU.
//   host : the seven random field elements (t, alpha, beta, gamma, delta, G1/G2 generator scalars) drawn exactly like
//          Fr::random_element; Lagrange evaluation u_i = L_i(t) over the basic / step domain with one batched inversion
//          (r1cs_to_qap.tcc:110-170, basic_radix2_domain_aux.tcc:183-235, step_radix2_domain.tcc:160-190); At/Bt/Ct by a sparse
//          transposed accumulation; scalar vectors of the A, B, H, L queries and gamma_ABC
//   GPU  : ~2 M G1 and ~0.3 M G2 fixed-base scalar multiplications (8-bit windows, 32 mixed additions each) + normalisation
// With the random words pinned the files are byte-identical to the reference's *_key output under the same word stream.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <random>
#include <string>
#include <vector>
#include "../../include/zkb200.h"
#include "prover.cuh"
#include "pk_format.hpp"
#include "msm.cuh"

using namespace zk;
using namespace zkh;
namespace zkv { void reduced_pairing(const HG1Affine &P, const HG2Affine &Q, HFq out[12]); }

namespace {
constexpr int FB_WINDOWS = 32;            // 8-bit windows: table of 32 x 255 multiples of the generator

// out[i] = scalar_i * G from the window table tab[w][d-1] = d * 2^(8w) * G; y_lsb[i] = parity of canonical y (for compression)
template <class F>
__global__ void __launch_bounds__(128) fixed_base_kernel(const Affine<F> *__restrict__ tab, const uint32_t *__restrict__ scalars, uint32_t n,
                                                         Affine<F> *__restrict__ out, uint8_t *__restrict__ y_lsb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s[8]; ld_scalar(scalars + (size_t)i * 8, s);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int w = 0; w < FB_WINDOWS; w++) {
        const uint32_t d = (s[w >> 2] >> ((w & 3) * 8)) & 255u;
        if (d) acc.add_affine(ld_affine(tab + (size_t)w * 255 + (d - 1)));
    }
    const Affine<F> a = acc.to_affine();
    out[i] = a;
    const uint32_t *yw = reinterpret_cast<const uint32_t *>(&a.y);          // first base-field element of y (c0 for G2)
    Fq y0; for (int k = 0; k < 8; k++) y0.v[k] = yw[k];
    y_lsb[i] = (uint8_t)(y0.from_mont().v[0] & 1);
}

template <class HP, class HA> static std::vector<HA> window_table(const HP &gen) {
    std::vector<HA> tab((size_t)FB_WINDOWS * 255);
    HP base = gen;
    for (int w = 0; w < FB_WINDOWS; w++) {
        HP cur = base;
        for (int d = 1; d <= 255; d++) { tab[(size_t)w * 255 + d - 1] = cur.to_affine(); cur = cur.add(base); }
        base = cur;                                   // 256 * base
    }
    return tab;
}

struct Compressed { std::vector<uint8_t> x; std::vector<uint8_t> flags; };     // x: 32/64 B Montgomery per point; flags bit0 = y lsb, bit1 = infinity
template <class F, class HA>
static Compressed fixed_base_batch(const std::vector<HA> &tab, const std::vector<HFr> &scalars) {
    const uint32_t n = (uint32_t)scalars.size();
    Compressed r; r.x.resize((size_t)n * sizeof(F)); r.flags.resize(n);
    if (!n) return r;
    std::vector<uint64_t> can((size_t)n * 4);
    for (uint32_t i = 0; i < n; i++) scalars[i].to_canonical(&can[(size_t)i * 4]);
    Affine<F> *d_tab, *d_out; uint32_t *d_sc; uint8_t *d_lsb;
    ZK_CUDA(cudaMalloc(&d_tab, tab.size() * sizeof(Affine<F>))); ZK_CUDA(cudaMalloc(&d_out, (size_t)n * sizeof(Affine<F>)));
    ZK_CUDA(cudaMalloc(&d_sc, (size_t)n * 32)); ZK_CUDA(cudaMalloc(&d_lsb, n));
    ZK_CUDA(cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(Affine<F>), cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMemcpy(d_sc, can.data(), (size_t)n * 32, cudaMemcpyHostToDevice));
    fixed_base_kernel<F><<<(n + 127) / 128, 128>>>(d_tab, d_sc, n, d_out, d_lsb);
    ZK_CUDA(cudaGetLastError());
    std::vector<Affine<F>> out(n); std::vector<uint8_t> lsb(n);
    ZK_CUDA(cudaMemcpy(out.data(), d_out, (size_t)n * sizeof(Affine<F>), cudaMemcpyDeviceToHost));
    ZK_CUDA(cudaMemcpy(lsb.data(), d_lsb, n, cudaMemcpyDeviceToHost));
    cudaFree(d_tab); cudaFree(d_out); cudaFree(d_sc); cudaFree(d_lsb);
    for (uint32_t i = 0; i < n; i++) {
        const bool inf = scalars[i].is_zero();
        if (inf) { memset(&r.x[(size_t)i * sizeof(F)], 0, sizeof(F)); r.flags[i] = 2; }
        else { memcpy(&r.x[(size_t)i * sizeof(F)], &out[i].x, sizeof(F)); r.flags[i] = lsb[i] & 1; }
    }
    return r;
}
static void put_point(std::string &o, const Compressed &c, size_t i, size_t xbytes) {
    // alt_bn128_g1.cpp:404-418 / alt_bn128_g2.cpp:418-431: zero flag, x (Montgomery, little-endian), y lsb -- '0'/'1' ASCII flags
    o.push_back((c.flags[i] & 2) ? '1' : '0');
    o.append((const char *)&c.x[i * xbytes], xbytes);
    o.push_back((c.flags[i] & 2) ? '1' : (char)('0' + (c.flags[i] & 1)));      // infinity prints the affine form (0, 1): lsb(y) = 1
}
static std::string dec_u64(uint64_t v) { return std::to_string(v); }
static std::string dec_fr(const HFr &x) {              // canonical decimal
    uint64_t c[4]; x.to_canonical(c);
    std::string out;
    while (c[0] | c[1] | c[2] | c[3]) {
        unsigned __int128 rem = 0;
        for (int i = 3; i >= 0; i--) { unsigned __int128 cur = (rem << 64) | c[i]; c[i] = (uint64_t)(cur / 10000000000000000000ull); rem = cur % 10000000000000000000ull; }
        uint64_t chunk = (uint64_t)rem;
        const bool last = !(c[0] | c[1] | c[2] | c[3]);
        for (int k = 0; k < 19 && (!last || chunk); k++) { out.push_back((char)('0' + chunk % 10)); chunk /= 10; }
    }
    if (out.empty()) out = "0";
    return std::string(out.rbegin(), out.rend());
}
static std::string dec_fq(const HFq &x) { HFr t; memcpy(t.v, x.v, 32); uint64_t c[4]; x.to_canonical(c);      // reuse the conversion on raw limbs
    HFr y = HFr::zero(); (void)t; (void)y;
    std::string out;
    while (c[0] | c[1] | c[2] | c[3]) {
        unsigned __int128 rem = 0;
        for (int i = 3; i >= 0; i--) { unsigned __int128 cur = (rem << 64) | c[i]; c[i] = (uint64_t)(cur / 10); rem = cur % 10; }
        out.push_back((char)('0' + (int)rem));
    }
    if (out.empty()) out = "0";
    return std::string(out.rbegin(), out.rend());
}

// Fr::random_element from a word stream (fp.tcc:695-721); returns the element (Montgomery form = the masked random integer)
struct WordStream {
    const uint32_t *w; size_t n, pos = 0; std::random_device rd;
    uint32_t next() { if (w) return w[pos++ % n]; return rd(); }
    HFr fr() {
        for (;;) {
            uint64_t m[4];
            for (int i = 0; i < 4; i++) { uint64_t lo = next(), hi = next(); m[i] = lo | (hi << 32); }
            m[3] &= 0x3fffffffffffffffull;
            if (!HFr::geq_mod(m)) return HFr::raw(m);
        }
    }
};
static HFr root_of_unity_h(int logn) {
    HFr w; const char *s = "19103219067921713944291392827692070036145651957329286315305642004821462161904";
    HFr::from_dec(s, strlen(s), w);
    for (int i = 28; i > logn; i--) w = w.sqr();
    return w;
}
static void batch_invert(std::vector<HFr> &v) {       // Montgomery's trick; all entries non-zero
    std::vector<HFr> pre(v.size());
    HFr acc = HFr::one();
    for (size_t i = 0; i < v.size(); i++) { pre[i] = acc; acc = acc * v[i]; }
    HFr inv = acc.inverse();
    for (size_t i = v.size(); i-- > 0;) { HFr t = inv * pre[i]; inv = inv * v[i]; v[i] = t; }
}
// _basic_radix2_evaluate_all_lagrange_polynomials (basic_radix2_domain_aux.tcc:183-235), t not in the domain
static std::vector<HFr> lagrange_basic(uint64_t m, int logm, const HFr &t) {
    if (m == 1) return std::vector<HFr>(1, HFr::one());
    const HFr omega = root_of_unity_h(logm);
    std::vector<HFr> den(m);
    HFr r = HFr::one();
    for (uint64_t i = 0; i < m; i++) { den[i] = t - r; r = r * omega; }
    batch_invert(den);
    HFr l = (t.pow64(m) - HFr::one()) * HFr::from_u64(m).inverse();
    for (uint64_t i = 0; i < m; i++) { den[i] = l * den[i]; l = l * omega; }
    return den;
}
static int ilog2c(uint64_t n) { int r = (n & (n - 1)) == 0 ? 0 : 1; while (n > 1) { n >>= 1; r++; } return r; }
} // namespace

int zkb200_keygen(const char *cs_source_pk_path, const uint32_t *words, size_t n_words, const char *out_pk_path, const char *out_vk_path,
                  double seconds[3]) {
    const auto T0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
    std::ifstream fh(cs_source_pk_path, std::ios::binary | std::ios::ate);
    if (!fh.is_open()) return -1;
    std::vector<uint8_t> data((size_t)fh.tellg());
    fh.seekg(0); fh.read((char *)data.data(), (std::streamsize)data.size());
    zkpk::ParsedPk P;
    if (!zkpk::parse_pk(data.data(), data.size(), P)) return -2;
    // the text of the constraint system is copied through verbatim (it is already in swap_AB_if_beneficial form, r1cs.tcc:182-231)
    size_t cs_off = 0;
    { zkpk::Cursor c(data.data(), data.size());
      // skip: 5 fixed points, A, B, H, L exactly as parse_pk does -- recompute the offset from the sizes
      cs_off = 34 + 1 + 34 + 1 + 66 + 1 + 34 + 1 + 66 + 1;
      cs_off += dec_u64(P.A.size()).size() + 1 + P.A.size() * 34;
      cs_off += dec_u64(P.B_domain).size() + 1 + dec_u64(P.B_idx.size()).size() + 1;
      for (uint32_t ix : P.B_idx) cs_off += dec_u64(ix).size() + 1;
      cs_off += dec_u64(P.B_idx.size()).size() + 1 + P.B_idx.size() * (66 + 1 + 34 + 1);
      cs_off += dec_u64(P.H.size()).size() + 1 + P.H.size() * 34;
      cs_off += dec_u64(P.L.size()).size() + 1 + P.L.size() * 34;
      (void)c; }
    const std::string cs_text((const char *)data.data() + cs_off, data.size() - cs_off);
    if (zkb200_ensure_device() < 0) return -3;

    const uint64_t ni = P.num_inputs, n = P.num_inputs + P.num_aux, nc = P.num_constraints;
    WordStream ws{n_words ? words : nullptr, n_words};
    const HFr t = ws.fr(), alpha = ws.fr(), beta = ws.fr(), gamma = ws.fr(), delta = ws.fr();
    const HFr gamma_inv = gamma.inverse(), delta_inv = delta.inverse();

    // domain of get_evaluation_domain(nc + ni + 1) and u = evaluate_all_lagrange_polynomials(t)
    const uint64_t min_size = nc + ni + 1;
    const int L = ilog2c(min_size);
    uint64_t m; bool step = false;
    if (min_size == (1ull << L)) m = min_size;
    else { const uint64_t big = 1ull << (L - 1), small = min_size - big, rs = 1ull << ilog2c(small);
           if (small == rs) { m = min_size; step = true; } else if (rs == big) m = 2 * big; else { m = big + rs; step = true; } }
    std::vector<HFr> u; HFr Zt;
    if (!step) { u = lagrange_basic(m, ilog2c(m), t); Zt = t.pow64(m) - HFr::one(); }
    else {
        // step_radix2_domain::evaluate_all_lagrange_polynomials / compute_vanishing_polynomial (step_radix2_domain.tcc:160-205)
        const uint64_t big = 1ull << (ilog2c(m) - 1), small = m - big;
        const HFr omega = root_of_unity_h(ilog2c(big) + 1), big_omega = omega.sqr();
        std::vector<HFr> ib = lagrange_basic(big, ilog2c(big), t), is = lagrange_basic(small, ilog2c(small), t * omega.inverse());
        const HFr o_sm = omega.pow64(small), L0 = t.pow64(small) - o_sm, step_elt = big_omega.pow64(small);
        std::vector<HFr> den(big);
        HFr elt = HFr::one();
        for (uint64_t i = 0; i < big; i++) { den[i] = elt - o_sm; elt = elt * step_elt; }
        batch_invert(den);
        u.resize(m);
        for (uint64_t i = 0; i < big; i++) u[i] = ib[i] * L0 * den[i];
        const HFr L1 = (t.pow64(big) - HFr::one()) * (omega.pow64(big) - HFr::one()).inverse();
        for (uint64_t i = 0; i < small; i++) u[big + i] = L1 * is[i];
        Zt = (t.pow64(big) - HFr::one()) * (t.pow64(small) - o_sm);
    }
    // At, Bt, Ct (r1cs_to_qap.tcc:121-147)
    std::vector<HFr> At(n + 1, HFr::zero()), Bt(n + 1, HFr::zero()), Ct(n + 1, HFr::zero());
    for (uint64_t i = 0; i <= ni; i++) At[i] = u[nc + i];
    const zkpk::Csr *mats[3] = {&P.a, &P.b, &P.c};
    std::vector<HFr> *outs[3] = {&At, &Bt, &Ct};
    for (int w = 0; w < 3; w++)
        for (uint64_t i = 0; i < nc; i++)
            for (uint32_t k = mats[w]->rowptr[i]; k < mats[w]->rowptr[i + 1]; k++) {
                const uint32_t ci = mats[w]->coef[k];
                HFr &dst = (*outs[w])[mats[w]->col[k]];
                dst = dst + (ci == 0 ? u[i] : u[i] * P.coef_dict[ci]);
            }
    // query scalars (r1cs_gg_ppzksnark.tcc:252-300)
    std::vector<HFr> gabc(ni + 1), Lt(n - ni), Hs(m - 1), Bnz; std::vector<uint32_t> Bidx;
    for (uint64_t i = 0; i <= ni; i++) gabc[i] = (beta * At[i] + alpha * Bt[i] + Ct[i]) * gamma_inv;
    for (uint64_t i = 0; i < n - ni; i++) Lt[i] = (beta * At[ni + 1 + i] + alpha * Bt[ni + 1 + i] + Ct[ni + 1 + i]) * delta_inv;
    { HFr ti = Zt * delta_inv; for (uint64_t i = 0; i + 1 < m; i++) { Hs[i] = ti; ti = ti * t; } }
    for (uint64_t i = 0; i <= n; i++) if (!Bt[i].is_zero()) { Bidx.push_back((uint32_t)i); Bnz.push_back(Bt[i]); }
    const HFr g1s = ws.fr(), g2s = ws.fr();                // G1::random_element, G2::random_element (r1cs_gg_ppzksnark.tcc:297,307)
    if (seconds) seconds[0] = since(T0);

    const auto T1 = std::chrono::steady_clock::now();
    uint64_t k1[4], k2[4]; g1s.to_canonical(k1); g2s.to_canonical(k2);
    const HG1 g1gen = HG1::from_affine(HG1Affine{HFq::one(), HFq::from_u64(2)}).mul(k1);
    HFq g2c[4]; const char *gs[4] = {"10857046999023057135944570762232829481370756359578518086990519993285655852781",
                                     "11559732032986387107991004021392285783925812861821192530917403151452391805634",
                                     "8495653923123431417604973247489272438418190587263600148770280649306958101930",
                                     "4082367875863433681332203403145435568316851327593401208105741076214120093531"};
    for (int i = 0; i < 4; i++) HFq::from_dec(gs[i], strlen(gs[i]), g2c[i]);
    const HG2 g2gen = HG2::from_affine(HG2Affine{HFq2{g2c[0], g2c[1]}, HFq2{g2c[2], g2c[3]}}).mul(k2);
    const std::vector<HG1Affine> tab1 = window_table<HG1, HG1Affine>(g1gen);
    const std::vector<HG2Affine> tab2 = window_table<HG2, HG2Affine>(g2gen);
    const std::vector<Affine<Fq>> &dt1 = reinterpret_cast<const std::vector<Affine<Fq>> &>(tab1);
    const std::vector<Affine<Fq2>> &dt2 = reinterpret_cast<const std::vector<Affine<Fq2>> &>(tab2);
    const Compressed fixed1 = fixed_base_batch<Fq>(dt1, {alpha, beta, delta});
    const Compressed fixed2 = fixed_base_batch<Fq2>(dt2, {beta, delta, gamma});
    const Compressed qA = fixed_base_batch<Fq>(dt1, At), qB1 = fixed_base_batch<Fq>(dt1, Bnz), qB2 = fixed_base_batch<Fq2>(dt2, Bnz);
    const Compressed qH = fixed_base_batch<Fq>(dt1, Hs), qL = fixed_base_batch<Fq>(dt1, Lt), qG = fixed_base_batch<Fq>(dt1, gabc);
    if (seconds) seconds[1] = since(T1);

    const auto T2 = std::chrono::steady_clock::now();
    // proving key (r1cs_gg_ppzksnark.tcc:52-66)
    std::string pk;
    pk.reserve(data.size() + 1024);
    put_point(pk, fixed1, 0, 32); pk.push_back('\n');      // alpha_g1
    put_point(pk, fixed1, 1, 32); pk.push_back('\n');      // beta_g1
    put_point(pk, fixed2, 0, 64); pk.push_back('\n');      // beta_g2
    put_point(pk, fixed1, 2, 32); pk.push_back('\n');      // delta_g1
    put_point(pk, fixed2, 1, 64); pk.push_back('\n');      // delta_g2
    pk += dec_u64(n + 1); pk.push_back('\n');
    for (uint64_t i = 0; i <= n; i++) put_point(pk, qA, i, 32);
    pk += dec_u64(n + 1); pk.push_back('\n'); pk += dec_u64(Bidx.size()); pk.push_back('\n');
    for (uint32_t ix : Bidx) { pk += dec_u64(ix); pk.push_back('\n'); }
    pk += dec_u64(Bidx.size()); pk.push_back('\n');
    for (size_t i = 0; i < Bidx.size(); i++) { put_point(pk, qB2, i, 64); pk.push_back(' '); put_point(pk, qB1, i, 32); pk.push_back('\n'); }
    pk += dec_u64(m - 1); pk.push_back('\n');
    for (uint64_t i = 0; i + 1 < m; i++) put_point(pk, qH, i, 32);
    pk += dec_u64(n - ni); pk.push_back('\n');
    for (uint64_t i = 0; i < n - ni; i++) put_point(pk, qL, i, 32);
    pk += cs_text;
    // verification key (r1cs_gg_ppzksnark.tcc:99-108)
    std::string vk;
    { // alpha_g1_beta_g2 = reduced_pairing(alpha*G1gen, beta*G2gen)
      uint64_t ka[4], kb[4]; alpha.to_canonical(ka); beta.to_canonical(kb);
      HFq gt[12]; zkv::reduced_pairing(g1gen.mul(ka).to_affine(), g2gen.mul(kb).to_affine(), gt);
      for (int i = 0; i < 12; i++) { vk += dec_fq(gt[i]); if (i < 11) vk.push_back(' '); }
      vk.push_back('\n'); }
    put_point(vk, fixed2, 2, 64); vk.push_back('\n');      // gamma_g2
    put_point(vk, fixed2, 1, 64); vk.push_back('\n');      // delta_g2
    put_point(vk, qG, 0, 32); vk.push_back('\n');          // gamma_ABC_g1.first, then the sparse_vector of the rest
    vk += dec_u64(ni); vk.push_back('\n'); vk += dec_u64(ni); vk.push_back('\n');
    for (uint64_t i = 0; i < ni; i++) { vk += dec_u64(i); vk.push_back('\n'); }
    vk += dec_u64(ni); vk.push_back('\n');
    for (uint64_t i = 0; i < ni; i++) { put_point(vk, qG, 1 + i, 32); vk.push_back('\n'); }
    vk += "\n\n";
    { std::ofstream o(out_pk_path, std::ios::binary); if (!o.is_open()) return -4; o.write(pk.data(), (std::streamsize)pk.size()); }
    { std::ofstream o(out_vk_path, std::ios::binary); if (!o.is_open()) return -4; o.write(vk.data(), (std::streamsize)vk.size()); }
    if (seconds) seconds[2] = since(T2);
    (void)dec_fr;
    return 0;
}
