// TEST INFRASTRUCTURE (oracle) -- not product code.
// Common C-ABI tail of the per-circuit harness libraries; included after the reference's <circuit>cgo.cpp.
// REF_NAME(x) pastes the circuit name, REF_CALL(pk) invokes the reference's generate_<circuit>_proof with
// the parsed arguments that the wrapper holds in scope.
#pragma once
namespace refhook {
static libsnark::r1cs_gg_ppzksnark_proving_key<ppT> g_pk;
static bool g_pk_loaded = false;
static int load_pk(const char *path) {
    ppT::init_public_params();
    g_pk = loadFromFile<libsnark::r1cs_gg_ppzksnark_proving_key<ppT>>(path);   // the reference's own loader (SRC/<c>/<c>cgo.cpp loadFromFile)
    g_pk_loaded = true;
    return 0;
}
// copy the recorded assignment out; returns number of variables
static long emit_witness(uint8_t *out, size_t cap_elems, int *satisfied) {
    State &S = g_state;
    if (satisfied) *satisfied = S.called ? 1 : 0;
    if (!S.called) return 0;
    size_t n = S.primary.size() + S.aux.size();
    if (out) {
        if (cap_elems < n) return -1;
        for (size_t i = 0; i < S.primary.size(); i++) put_fp(out + 32 * i, S.primary[i]);
        for (size_t i = 0; i < S.aux.size(); i++) put_fp(out + 32 * (S.primary.size() + i), S.aux[i]);
    }
    return (long)n;
}
static int begin(int mode, const uint32_t *words, size_t nwords, const char *outdir) {
    ppT::init_public_params();
    State &S = g_state;
    S.mode = mode; S.called = false; S.outdir = outdir ? outdir : "";
    if (mode == 1) {
        if (!g_pk_loaded) return -2;
        FrT rs[2];
        if (fr_from_words(words, nwords, rs, 2) < 0) return -3;
        S.r = rs[0]; S.s = rs[1];
    }
    return 0;
}
static int finish_prove(const libsnark::r1cs_gg_ppzksnark_proof<ppT> &proof, char *proof_hex, uint8_t *pts, double *timings) {
    State &S = g_state;
    std::string h = string_proof_as_hex(proof);       // the reference's own encoder (SRC/<c>/<c>cgo.cpp:176-187)
    if (proof_hex) { memcpy(proof_hex, h.data(), 512); proof_hex[512] = 0; }
    if (!S.called) return 1;                           // unsatisfied: default proof, prover never ran
    if (pts) memcpy(pts, S.pts, sizeof(S.pts));
    if (timings) { timings[0] = S.t_qap; timings[1] = S.t_A; timings[2] = S.t_B; timings[3] = S.t_H; timings[4] = S.t_L; }
    if (!S.outdir.empty()) write_file(S.outdir + "/proof.hex", h.data(), 512);
    return 0;
}
// ---- verification with an explicit key path -------------------------------------------------------------------------------------
// The reference's verify<Circuit>proof (SRC/<c>/<c>cgo.cpp) reads /usr/local/prfKey/<c>vk.txt, a path compiled in.  ref_<c>_verify below
// runs the same steps on a key file given by the caller: the reference's own loader (deserializevkFromFile), the proof decoding of
// verify<Circuit>proof (mintcgo.cpp:340-404: hex pairs -> bytes by convertFromAscii, big-endian bytes -> bigint by libsnarkBigintFromBytes,
// bigint assigned to the coordinate, i.e. Fp_model(bigint) which REDUCES mod q; Z stays one), then the reference's verify_<c>_proof
// template, which packs the public inputs (gadget::witness_map) and calls r1cs_gg_ppzksnark_verifier_strong_IC (r1cs_gg_ppzksnark.tcc:613-623).
static libsnark::r1cs_gg_ppzksnark_verification_key<ppT> g_vk;
static std::string g_vk_path;
static const libsnark::r1cs_gg_ppzksnark_verification_key<ppT> &load_vk(const char *path) {
    ppT::init_public_params();
    if (g_vk_path != path) { g_vk = deserializevkFromFile(path); g_vk_path = path; }
    return g_vk;
}
static libff::bigint<libff::alt_bn128_r_limbs> coord_from_hex(const char *hex64) {
    uint8_t raw[64];
    for (int i = 0, j = 0; i < 64; i += 2, j++) raw[j] = uint8_t(convertFromAscii(uint8_t(hex64[i])) * 16 + convertFromAscii(uint8_t(hex64[i + 1])));
    return libsnarkBigintFromBytes(raw);
}
static libsnark::r1cs_gg_ppzksnark_proof<ppT> proof_from_hex(const char *data) {
    libsnark::r1cs_gg_ppzksnark_proof<ppT> proof;          // default: the three generators, so every Z is one
    proof.g_A.X = coord_from_hex(data); proof.g_A.Y = coord_from_hex(data + 64);
    proof.g_B.X.c1 = coord_from_hex(data + 128); proof.g_B.X.c0 = coord_from_hex(data + 192);
    proof.g_B.Y.c1 = coord_from_hex(data + 256); proof.g_B.Y.c0 = coord_from_hex(data + 320);
    proof.g_C.X = coord_from_hex(data + 384); proof.g_C.Y = coord_from_hex(data + 448);
    return proof;
}
} // namespace refhook
