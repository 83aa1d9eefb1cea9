// TEST INFRASTRUCTURE (oracle) -- not product code.
// Included by ref_{mint,send,deposit,redeem}.cpp BEFORE they #include the reference's own
// <circuit>cgo.cpp (compiled in place from /root/reference, never copied).  It defines a prover hook and then
// renames the prover symbol with a macro, so that the reference's generate_<circuit>_proof() template
// (SRC/<c>/<c>cgo.cpp) -- gadget construction, constraint generation, witness generation, is_satisfied --
// runs unmodified and hands (pk, primary_input, auxiliary_input) to ref_hooked_prover() below instead of
// r1cs_gg_ppzksnark_prover().  The hook either
//   mode 0: records the full assignment only (no pk needed), or
//   mode 1: runs the reference prover pipeline phase by phase with PINNED r,s and dumps every intermediate:
//           the same library calls in the same order as r1cs_gg_ppzksnark.tcc:402-501.
#pragma once
#include <boost/optional.hpp>
#include <boost/foreach.hpp>
#include <boost/format.hpp>
#include <sys/time.h>
#include "libsnark/zk_proof_systems/ppzksnark/r1cs_gg_ppzksnark/r1cs_gg_ppzksnark.hpp"
#include "libsnark/common/default_types/r1cs_gg_ppzksnark_pp.hpp"
#include <libsnark/gadgetlib1/gadgets/hashes/sha256/sha256_gadget.hpp>
#include "libff/algebra/curves/alt_bn128/alt_bn128_pp.hpp"
#ifdef MULTICORE
#include <omp.h>
#endif
#include "ref_io.h"

namespace refhook {
using namespace refio;
struct State {
    int mode = 0;                 // 0 = witness only, 1 = pinned prove + dump
    bool called = false;
    std::vector<FrT> primary, aux;
    FrT r, s;                     // pinned prover randomness (mode 1)
    std::string outdir;           // where to dump (mode 1), may be empty
    std::string proof_hex;        // filled by the circuit wrapper
    double t_qap = 0, t_A = 0, t_B = 0, t_H = 0, t_L = 0;
    uint8_t pts[64 + 128 + 64 + 64 + 64]; // At, Bt.g, Bt.h, Ht, Lt
    size_t m = 0;
};
static State g_state;
static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

// Turn 16 words of a std::random_device-style stream into (r, s) exactly as Fr::random_element does
// (FF/algebra/fields/fp.tcc:695-721 + bigint.tcc:167-179): 8 words -> mont_repr (word 0 = low 32 bits of limb 0),
// clear bits >= 254, reject if >= modulus.  Returns number of words consumed, or -1 if the stream ran out.
static int fr_from_words(const uint32_t *w, size_t nwords, FrT *out, int count) {
    size_t pos = 0;
    for (int k = 0; k < count; k++) {
        for (;;) {
            if (pos + 8 > nwords) return -1;
            FrT x;
            memcpy(x.mont_repr.data, w + pos, 32); pos += 8;
            x.mont_repr.data[3] &= 0x3fffffffffffffffull;
            if (mpn_cmp(x.mont_repr.data, FrT::mod.data, 4) < 0) { out[k] = x; break; }
        }
    }
    return (int)pos;
}
} // namespace refhook

template <typename ppT>
libsnark::r1cs_gg_ppzksnark_proof<ppT> ref_hooked_prover(const libsnark::r1cs_gg_ppzksnark_proving_key<ppT> &pk,
                                                         const libsnark::r1cs_gg_ppzksnark_primary_input<ppT> &primary_input,
                                                         const libsnark::r1cs_gg_ppzksnark_auxiliary_input<ppT> &auxiliary_input) {
    using namespace refhook;
    using namespace libsnark;
    State &S = g_state;
    S.called = true;
    S.primary = primary_input; S.aux = auxiliary_input;
    if (S.mode == 0) return r1cs_gg_ppzksnark_proof<ppT>();

    double t = now();
    const qap_witness<FrT> w = r1cs_to_qap_witness_map(pk.constraint_system, primary_input, auxiliary_input, FrT::zero(), FrT::zero(), FrT::zero());
    S.t_qap = now() - t; S.m = w.degree();
#ifdef MULTICORE
    const size_t chunks = omp_get_max_threads();
#else
    const size_t chunks = 1;
#endif
    const FrT r = S.r, s = S.s;
    libff::Fr_vector<ppT> cpa(1, FrT::one());
    cpa.insert(cpa.end(), w.coefficients_for_ABCs.begin(), w.coefficients_for_ABCs.end());
    t = now();
    G1T eA = libff::multi_exp_with_mixed_addition<G1T, FrT, libff::multi_exp_method_BDLO12>(
        pk.A_query.begin(), pk.A_query.begin() + w.num_variables() + 1, cpa.begin(), cpa.begin() + w.num_variables() + 1, chunks);
    S.t_A = now() - t; t = now();
    auto eB = kc_multi_exp_with_mixed_addition<G2T, G1T, FrT, libff::multi_exp_method_BDLO12>(
        pk.B_query, 0, w.num_variables() + 1, cpa.begin(), cpa.begin() + w.num_variables() + 1, chunks);
    S.t_B = now() - t; t = now();
    G1T eH = libff::multi_exp<G1T, FrT, libff::multi_exp_method_BDLO12>(
        pk.H_query.begin(), pk.H_query.begin() + (w.degree() - 1), w.coefficients_for_H.begin(), w.coefficients_for_H.begin() + (w.degree() - 1), chunks);
    S.t_H = now() - t; t = now();
    G1T eL = libff::multi_exp_with_mixed_addition<G1T, FrT, libff::multi_exp_method_BDLO12>(
        pk.L_query.begin(), pk.L_query.end(), cpa.begin() + w.num_inputs() + 1, cpa.begin() + w.num_variables() + 1, chunks);
    S.t_L = now() - t;

    G1T g1_A = pk.alpha_g1 + eA + r * pk.delta_g1;
    G1T g1_B = pk.beta_g1 + eB.h + s * pk.delta_g1;
    G2T g2_B = pk.beta_g2 + eB.g + s * pk.delta_g2;
    G1T g1_C = eH + eL + s * g1_A + r * g1_B - (r * s) * pk.delta_g1;

    put_g1(S.pts, eA); put_g2(S.pts + 64, eB.g); put_g1(S.pts + 192, eB.h); put_g1(S.pts + 256, eH); put_g1(S.pts + 320, eL);
    if (!S.outdir.empty()) {
        std::vector<FrT> full(primary_input); full.insert(full.end(), auxiliary_input.begin(), auxiliary_input.end());
        write_fr_vec(S.outdir + "/assignment.bin", full, full.size());
        write_fr_vec(S.outdir + "/coefficients_for_H.bin", w.coefficients_for_H, w.coefficients_for_H.size());
        write_file(S.outdir + "/msm_points.bin", S.pts, sizeof(S.pts));
        uint8_t rs[64]; put_fp(rs, r); put_fp(rs + 32, s);
        write_file(S.outdir + "/rs.bin", rs, 64);
    }
    return r1cs_gg_ppzksnark_proof<ppT>(std::move(g1_A), std::move(g2_B), std::move(g1_C));
}

// From here on, every textual use of the prover's name in the reference TU resolves to the hook.
#define r1cs_gg_ppzksnark_prover ref_hooked_prover
