// TEST INFRASTRUCTURE (oracle) -- not product code.  Harness around the reference's mint circuit TU.
#include "ref_circuit.h"
#include "mint/mintcgo.cpp"          // reference TU, compiled in place (-I /root/reference/libsnark-vnt/src)
#include "ref_circuit_tail.h"
using namespace refhook;

// argument parsing as genMintproof does it (SRC/mint/mintcgo.cpp:280-291), then the reference's generate_mint_proof
static libsnark::r1cs_gg_ppzksnark_proof<ppT> run(uint64_t value, uint64_t value_old, const char *sn_old_s, const char *r_old_s,
        const char *sn_s, const char *r_s, const char *cmtA_old_s, const char *cmtA_s, uint64_t value_s, const char *sk_s) {
    uint256 sn_old = uint256S(sn_old_s), r_old = uint256S(r_old_s), sn = uint256S(sn_s), r = uint256S(r_s);
    uint256 cmtA_old = uint256S(cmtA_old_s), cmtA = uint256S(cmtA_s), sk = uint256S(sk_s);
    Note note_old = Note(value_old, sn_old, r_old);
    Note note = Note(value, sn, r);
    return generate_mint_proof<ppT>(g_state.mode == 1 ? g_pk : libsnark::r1cs_gg_ppzksnark_proving_key<ppT>(),
                                    note_old, note, cmtA_old, cmtA, value_s, sk);
}
extern "C" {
int ref_mint_load_pk(const char *path) { return load_pk(path); }
long ref_mint_witness(uint64_t value, uint64_t value_old, const char *sn_old, const char *r_old, const char *sn, const char *r,
                      const char *cmtA_old, const char *cmtA, uint64_t value_s, const char *sk,
                      uint8_t *out, size_t cap_elems, int *satisfied) {
    if (begin(0, nullptr, 0, nullptr)) return -1;
    run(value, value_old, sn_old, r_old, sn, r, cmtA_old, cmtA, value_s, sk);
    return emit_witness(out, cap_elems, satisfied);
}
int ref_mint_prove(uint64_t value, uint64_t value_old, const char *sn_old, const char *r_old, const char *sn, const char *r,
                   const char *cmtA_old, const char *cmtA, uint64_t value_s, const char *sk,
                   const uint32_t *words, size_t nwords, const char *outdir, char *proof_hex, uint8_t *pts, double *timings) {
    int rc = begin(1, words, nwords, outdir); if (rc) return rc;
    auto proof = run(value, value_old, sn_old, r_old, sn, r, cmtA_old, cmtA, value_s, sk);
    return finish_prove(proof, proof_hex, pts, timings);
}
int ref_mint_verify(const char *vk_path, const char *proof_hex, const char *cmtA_old_s, const char *sn_old_s, const char *cmtA_s, uint64_t value_s) {
    uint256 sn_old = uint256S(sn_old_s), cmtA_old = uint256S(cmtA_old_s), cmtA = uint256S(cmtA_s);
    const auto &vk = load_vk(vk_path);            // first: it also runs init_public_params(), which the proof's default constructor needs
    return verify_mint_proof<ppT>(vk, proof_from_hex(proof_hex), cmtA_old, sn_old, cmtA, value_s) ? 1 : 0;
}
}
