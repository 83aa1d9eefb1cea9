// TEST INFRASTRUCTURE (oracle) -- not product code.  Harness around the reference's deposit circuit TU.
#include "ref_circuit.h"
#include "deposit/depositcgo.cpp"    // reference TU, compiled in place
#include "ref_circuit_tail.h"
using namespace refhook;

// argument parsing + Merkle witness as genDepositproof does it (SRC/deposit/depositcgo.cpp:352-403), then generate_deposit_proof
static libsnark::r1cs_gg_ppzksnark_proof<ppT> run(uint64_t value, uint64_t value_old, const char *sn_old_s, const char *r_old_s, const char *sn_str,
        const char *r_str, const char *sns_s, const char *rs_s, const char *cmtB_old_s, const char *cmtB_s, uint64_t value_s, const char *pk_s,
        const char *sn_A_old_s, const char *cmtS_s, const char *cmtarray, int n, const char *sk_s) {
    uint256 sn_old = uint256S(sn_old_s), r_old = uint256S(r_old_s), sn = uint256S(sn_str), r = uint256S(r_str);
    uint256 sn_s = uint256S(sns_s), r_s = uint256S(rs_s), cmtB_old = uint256S(cmtB_old_s), cmtB = uint256S(cmtB_s);
    uint160 pk_recv = uint160S(pk_s);
    uint256 sn_A_old = uint256S(sn_A_old_s), cmtS = uint256S(cmtS_s), sk = uint256S(sk_s);
    Note note_old = Note(value_old, sn_old, r_old);
    NoteS note_s = NoteS(value_s, pk_recv, r_s, sn_A_old);
    Note note = Note(value, sn, r);
    std::vector<uint256> commitments(n);
    std::string all = cmtarray;
    for (int i = 0; i < n; i++) commitments[i] = uint256S(all.substr(i * 66, 66));
    ZCIncrementalMerkleTree tree;
    ZCIncrementalWitness wit = tree.witness();
    bool found = false;
    for (int i = 0; i < n; i++) {
        if (found) wit.append(commitments[i]); else tree.append(commitments[i]);
        if (commitments[i] == cmtS) { wit = tree.witness(); found = true; }
    }
    auto path = wit.path();
    uint256 rt = wit.root();
    return generate_deposit_proof<ppT>(g_state.mode == 1 ? g_pk : libsnark::r1cs_gg_ppzksnark_proving_key<ppT>(),
                                       note_s, note_old, note, cmtS, cmtB_old, cmtB, rt, path, sn_s, sk);
}
extern "C" {
int ref_deposit_load_pk(const char *path) { return load_pk(path); }
long ref_deposit_witness(uint64_t value, uint64_t value_old, const char *sn_old, const char *r_old, const char *sn, const char *r,
                         const char *sns, const char *rs, const char *cmtB_old, const char *cmtB, uint64_t value_s, const char *pk,
                         const char *sn_A_old, const char *cmtS, const char *cmtarray, int n, const char *RT, const char *sk,
                         uint8_t *out, size_t cap_elems, int *satisfied) {
    (void)RT;
    if (begin(0, nullptr, 0, nullptr)) return -1;
    run(value, value_old, sn_old, r_old, sn, r, sns, rs, cmtB_old, cmtB, value_s, pk, sn_A_old, cmtS, cmtarray, n, sk);
    return emit_witness(out, cap_elems, satisfied);
}
int ref_deposit_prove(uint64_t value, uint64_t value_old, const char *sn_old, const char *r_old, const char *sn, const char *r,
                      const char *sns, const char *rs, const char *cmtB_old, const char *cmtB, uint64_t value_s, const char *pk,
                      const char *sn_A_old, const char *cmtS, const char *cmtarray, int n, const char *RT, const char *sk,
                      const uint32_t *words, size_t nwords, const char *outdir, char *proof_hex, uint8_t *pts, double *timings) {
    (void)RT;
    int rc = begin(1, words, nwords, outdir); if (rc) return rc;
    auto proof = run(value, value_old, sn_old, r_old, sn, r, sns, rs, cmtB_old, cmtB, value_s, pk, sn_A_old, cmtS, cmtarray, n, sk);
    return finish_prove(proof, proof_hex, pts, timings);
}
int ref_deposit_verify(const char *vk_path, const char *proof_hex, const char *RT, const char *pk, const char *cmtb_old, const char *snold, const char *cmtb,
                       const char *sns) {
    uint256 rt = uint256S(RT), cmtB_old = uint256S(cmtb_old), sn_old = uint256S(snold), cmtB = uint256S(cmtb), sn_s = uint256S(sns);
    uint160 pk_recv = uint160S(pk);
    const auto &vk = load_vk(vk_path);            // first: it also runs init_public_params(), which the proof's default constructor needs
    return verify_deposit_proof<ppT>(vk, proof_from_hex(proof_hex), rt, pk_recv, cmtB_old, sn_old, cmtB, sn_s) ? 1 : 0;
}
}
