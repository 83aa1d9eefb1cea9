// Host-side "witness handling" for BlockMaze's four circuits: byte-level helpers of the cgo layer (hex blobs, SHA-256
// commitments, the depth-8 incremental Merkle tree) and native generators of the FULL variable assignment that the
// reference obtains by running its gadgetlib1 circuits (protoboard::full_variable_assignment).
//
// The assignment layout is an artefact of gadget construction order in the reference (protoboard::allocate_var_index,
// libsnark/gadgetlib1/protoboard.tcc:37-49); each generator below documents the order it reproduces.  Parity is checked
// element-for-element against the reference gadgets (tests/test_host.py via oracle/_ref).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "witness_sha.hpp"

namespace zkw {

// ---- byte helpers (uint256.h base_blob::SetHex/GetHex, util.h, Note.h) -------------------------------------------------
// SetHex: skips leading whitespace and "0x", reads hex digits from the END of the string into data[0..]  (uint256.h:200-226)
void parse_hex_blob(const char *s, uint8_t *out, size_t nbytes);
std::string blob_to_hex(const uint8_t *data, size_t nbytes);          // GetHex: reversed bytes as lowercase hex
void sha256(const uint8_t *data, size_t len, uint8_t out[32]);        // CSHA256 Write+Finalize (standard padding)
void sha256_compress(const uint8_t left[32], const uint8_t right[32], uint8_t out[32]);   // FinalizeNoPadding: one block from the IV
void note_cm(uint64_t value, const uint8_t sn[32], const uint8_t r[32], uint8_t out[32]);                          // Note::cm
void notes_cm(uint64_t value, const uint8_t pk[20], const uint8_t r[32], const uint8_t sn_old[32], uint8_t out[32]); // NoteS::cm
void compute_prf(const uint8_t sk[32], const uint8_t r[32], uint8_t out[32]);                                      // Compute_PRF
void compute_crh(const uint8_t pk[20], const uint8_t r[32], uint8_t out[32]);                                      // Compute_CRH

// Depth-8 tree over `n` leaves padded with zero leaves (ZCIncrementalMerkleTree::root()).
constexpr int MERKLE_DEPTH = 8;      // INCREMENTAL_MERKLE_TREE_DEPTH (deposit/VNT.h:6)
void merkle_root(const uint8_t (*leaves)[32], size_t n, uint8_t out[32]);
// Authentication path of leaf `index`: siblings[d] for level d counted from the LEAF (d = 0) up; returns the root.
void merkle_path(const uint8_t (*leaves)[32], size_t n, size_t index, uint8_t siblings[MERKLE_DEPTH][32], uint8_t root[32]);

// ---- full assignments -----------------------------------------------------------------------------------------------------
// Compact form: ~97 % of a BlockMaze assignment is 0/1 and all but a handful of values fit 64 bits, so the assignment is kept as one
// uint64 per variable (`lo`) plus a short list of values wider than 64 bits.  This is what travels to the GPU (8 B instead of 32 B
// per variable); expand() produces the canonical 32-byte form of the C-ABI.
struct WideValue { uint32_t idx; uint32_t pad; uint64_t v[4]; };      // idx = variable index (1-based, 0 is the constant ONE)
// One sha256_compression_function_gadget whose 24 792 internal variables (a contiguous run starting at `base`) are left for the GPU to
// fill in (prover.cu sha256_witness_kernel): all of them are bits and sums of the 64-round trace of ONE compression, which is determined
// by the 16 message words and the 8 words of the incoming chaining value.  The host still evaluates the compression itself (microseconds) for
// the digest bits that live outside the run.
struct CompressionSeed { uint32_t base; uint32_t w[16]; uint32_t h[8]; };
constexpr uint32_t COMPRESSION_VARS = SHA_RUN_VARS;
struct Assignment {
    std::vector<uint64_t> own;        // (num_vars + 1) low words, variable i at [i], variable 0 = constant ONE
    uint64_t *ext = nullptr;          // caller-provided storage of the same shape (e.g. pinned host memory) used instead of `own`
    std::vector<WideValue> wide;
    std::vector<CompressionSeed> seeds;   // deferred compressions (empty unless the generator was asked to defer): their runs in lo() are unset
    uint32_t num_vars = 0;
    const uint64_t *lo() const { return ext ? ext : own.data(); }
    void expand(uint8_t *out) const {                                  // num_vars x 32 B canonical little-endian
        const uint64_t *l = lo();
        uint64_t *o = reinterpret_cast<uint64_t *>(out);
        for (uint32_t i = 1; i <= num_vars; i++) { o[4 * (i - 1)] = l[i]; o[4 * (i - 1) + 1] = o[4 * (i - 1) + 2] = o[4 * (i - 1) + 3] = 0; }
        for (const WideValue &w : wide) memcpy(o + 4 * (size_t)(w.idx - 1), w.v, 32);
        for (const CompressionSeed &sd : seeds) {                      // deferred runs: the same expansion the GPU kernel does
            ShaTrace T; sha_trace_build(sd.w, sd.h, T);
            for (uint32_t v = 0; v < SHA_RUN_VARS; v++) o[4 * (size_t)(sd.base + v - 1)] = sha_trace_value(T, v);
        }
    }
};
// Every generator takes an optional `ext` buffer of (num_vars + 1) uint64 to build the low words in place, and `defer`: leave the runs of
// the SHA-256 compression gadgets (97 % of the variables) to the GPU and return their seeds instead (Assignment::seeds).

struct Note { uint64_t value; uint8_t sn[32]; uint8_t r[32]; };
struct NoteS { uint64_t value; uint8_t pk[20]; uint8_t r[32]; uint8_t sn_old[32]; };

// mint_gadget::generate_r1cs_witness (SRC/mint/circuit/gadget.tcc:194-246)
Assignment mint_witness(const Note &note_old, const Note &note, const uint8_t cmtA_old[32], const uint8_t cmtA[32], uint64_t value_s,
                        const uint8_t sk[32], uint64_t *ext = nullptr, bool defer = false);
// redeem_gadget::generate_r1cs_witness (SRC/redeem/circuit/gadget.tcc)
Assignment redeem_witness(const Note &note_old, const Note &note, const uint8_t cmtA_old[32], const uint8_t cmtA[32], uint64_t value_s,
                          const uint8_t sk[32], uint64_t *ext = nullptr, bool defer = false);
// send_gadget::generate_r1cs_witness (SRC/send/circuit/gadget.tcc)
Assignment send_witness(const Note &note_old, const NoteS &note_s, const Note &note, const uint8_t cmtA_old[32], const uint8_t cmtS[32],
                        const uint8_t cmtA[32], const uint8_t sk[32], const uint8_t pk_sender[20], uint64_t *ext = nullptr, bool defer = false);
// deposit_gadget::generate_r1cs_witness (SRC/deposit/circuit/gadget.tcc); path as produced by merkle_path()
Assignment deposit_witness(const NoteS &note_s, const Note &note_old, const Note &note, const uint8_t cmtS[32], const uint8_t cmtB_old[32],
                           const uint8_t cmtB[32], const uint8_t rt[32], size_t leaf_index, const uint8_t siblings[MERKLE_DEPTH][32],
                           const uint8_t sn_s[32], const uint8_t sk[32], uint64_t *ext = nullptr, bool defer = false);

// directory of <circuit>{pk,vk}.txt: zkb200_set_key_dir() > $ZKB200_KEY_DIR > /usr/local/prfKey (defined in blockmaze_abi.cu)
std::string key_dir();

constexpr uint32_t MINT_VARS = 151512, SEND_VARS = 227046, DEPOSIT_VARS = 457127, REDEEM_VARS = 151579;

} // namespace zkw
