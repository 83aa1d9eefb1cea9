// TEST INFRASTRUCTURE (oracle) -- not product code.  Harness around the reference's send circuit TU.
#include "ref_circuit.h"
#include "send/sendcgo.cpp"          // reference TU, compiled in place
#include "ref_circuit_tail.h"
using namespace refhook;

// argument parsing as genSendproof does it (SRC/send/sendcgo.cpp:318-334,354), then the reference's generate_send_proof
static libsnark::r1cs_gg_ppzksnark_proof<ppT> run(uint64_t value_A, const char *r_s_s, const char *sn_s, const char *r_str, const char *cmt_s_s,
        const char *cmtA_s, uint64_t value_s, const char *pk_recv_s, uint64_t value_A_new, const char *sn_A_new, const char *r_A_new,
        const char *cmt_A_new, const char *sk_s, const char *pk_sender_s) {
    uint256 r_s = uint256S(r_s_s), sn = uint256S(sn_s), r = uint256S(r_str), cmtS = uint256S(cmt_s_s), cmtA = uint256S(cmtA_s);
    uint160 pk_recv = uint160S(pk_recv_s);
    uint256 snAnew = uint256S(sn_A_new), rAnew = uint256S(r_A_new), cmtAnew = uint256S(cmt_A_new), sk = uint256S(sk_s);
    uint160 pk_sender = uint160S(pk_sender_s);
    Note note_old = Note(value_A, sn, r);
    NoteS notes = NoteS(value_s, pk_recv, r_s, sn);
    Note note_new = Note(value_A_new, snAnew, rAnew);
    return generate_send_proof<ppT>(g_state.mode == 1 ? g_pk : libsnark::r1cs_gg_ppzksnark_proving_key<ppT>(),
                                    note_old, notes, note_new, cmtA, cmtS, cmtAnew, sk, pk_sender);
}
extern "C" {
int ref_send_load_pk(const char *path) { return load_pk(path); }
long ref_send_witness(uint64_t value_A, const char *r_s, const char *sn, const char *r, const char *cmt_s, const char *cmtA, uint64_t value_s,
                      const char *pk_recv, uint64_t value_A_new, const char *sn_A_new, const char *r_A_new, const char *cmt_A_new,
                      const char *sk, const char *pk_sender, uint8_t *out, size_t cap_elems, int *satisfied) {
    if (begin(0, nullptr, 0, nullptr)) return -1;
    run(value_A, r_s, sn, r, cmt_s, cmtA, value_s, pk_recv, value_A_new, sn_A_new, r_A_new, cmt_A_new, sk, pk_sender);
    return emit_witness(out, cap_elems, satisfied);
}
int ref_send_prove(uint64_t value_A, const char *r_s, const char *sn, const char *r, const char *cmt_s, const char *cmtA, uint64_t value_s,
                   const char *pk_recv, uint64_t value_A_new, const char *sn_A_new, const char *r_A_new, const char *cmt_A_new,
                   const char *sk, const char *pk_sender,
                   const uint32_t *words, size_t nwords, const char *outdir, char *proof_hex, uint8_t *pts, double *timings) {
    int rc = begin(1, words, nwords, outdir); if (rc) return rc;
    auto proof = run(value_A, r_s, sn, r, cmt_s, cmtA, value_s, pk_recv, value_A_new, sn_A_new, r_A_new, cmt_A_new, sk, pk_sender);
    return finish_prove(proof, proof_hex, pts, timings);
}
int ref_send_verify(const char *vk_path, const char *proof_hex, const char *cmtA_old_s, const char *sn_old_s, const char *cmtS_s, const char *cmtA_new_s) {
    uint256 cmtA_old = uint256S(cmtA_old_s), sn_old = uint256S(sn_old_s), cmtS = uint256S(cmtS_s), cmtA_new = uint256S(cmtA_new_s);
    const auto &vk = load_vk(vk_path);            // first: it also runs init_public_params(), which the proof's default constructor needs
    return verify_send_proof<ppT>(vk, proof_from_hex(proof_hex), cmtA_old, sn_old, cmtS, cmtA_new) ? 1 : 0;
}
}
