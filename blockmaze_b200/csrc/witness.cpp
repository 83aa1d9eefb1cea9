// See witness.hpp.  Everything here is integer / bit arithmetic on the host; the heavy lifting (SpMV, NTT, MSM) that
// consumes the assignment runs on the GPU.
#include "witness.hpp"
#include "host_field.hpp"

namespace zkw {

// =====================================================================================================================
// SHA-256 (FIPS 180-4).  CSHA256 in the reference is Bitcoin's implementation (SRC/*/deps/sha256.h).
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static const uint32_t IV256[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static void sha256_block(uint32_t st[8], const uint8_t blk[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K256[i] + w[i];
        uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
static void put_state(const uint32_t st[8], uint8_t out[32]) {
    for (int i = 0; i < 8; i++) { out[4 * i] = st[i] >> 24; out[4 * i + 1] = st[i] >> 16; out[4 * i + 2] = st[i] >> 8; out[4 * i + 3] = st[i]; }
}
void sha256(const uint8_t *data, size_t len, uint8_t out[32]) {
    uint32_t st[8]; memcpy(st, IV256, 32);
    size_t off = 0;
    for (; off + 64 <= len; off += 64) sha256_block(st, data + off);
    uint8_t tail[128]; memset(tail, 0, sizeof tail);
    const size_t rem = len - off;
    memcpy(tail, data + off, rem);
    tail[rem] = 0x80;
    const size_t tl = rem + 9 <= 64 ? 64 : 128;
    const uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_block(st, tail);
    if (tl == 128) sha256_block(st, tail + 64);
    put_state(st, out);
}
void sha256_compress(const uint8_t left[32], const uint8_t right[32], uint8_t out[32]) {
    uint32_t st[8]; memcpy(st, IV256, 32);
    uint8_t blk[64]; memcpy(blk, left, 32); memcpy(blk + 32, right, 32);
    sha256_block(st, blk);
    put_state(st, out);
}

// =====================================================================================================================
static int hexval(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}
void parse_hex_blob(const char *s, uint8_t *out, size_t nbytes) {
    memset(out, 0, nbytes);
    if (!s) return;
    while (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r' || *s == '\f' || *s == '\v') s++;
    if (s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) s += 2;
    const char *b = s;
    while (hexval(*s) >= 0) s++;
    s--;
    uint8_t *p = out, *end = out + nbytes;
    while (s >= b && p < end) {
        *p = (uint8_t)hexval(*s--);
        if (s >= b) { *p |= (uint8_t)(hexval(*s--) << 4); p++; }
    }
}
std::string blob_to_hex(const uint8_t *data, size_t nbytes) {
    static const char *d = "0123456789abcdef";
    std::string o(nbytes * 2, '0');
    for (size_t i = 0; i < nbytes; i++) { o[2 * i] = d[data[nbytes - 1 - i] >> 4]; o[2 * i + 1] = d[data[nbytes - 1 - i] & 15]; }
    return o;
}
static void put_le64(uint8_t *o, uint64_t v) { for (int i = 0; i < 8; i++) o[i] = (uint8_t)(v >> (8 * i)); }
void note_cm(uint64_t value, const uint8_t sn[32], const uint8_t r[32], uint8_t out[32]) {
    uint8_t m[72]; put_le64(m, value); memcpy(m + 8, sn, 32); memcpy(m + 40, r, 32); sha256(m, 72, out);
}
void notes_cm(uint64_t value, const uint8_t pk[20], const uint8_t r[32], const uint8_t sn_old[32], uint8_t out[32]) {
    uint8_t m[92]; put_le64(m, value); memcpy(m + 8, pk, 20); memcpy(m + 28, r, 32); memcpy(m + 60, sn_old, 32); sha256(m, 92, out);
}
void compute_prf(const uint8_t sk[32], const uint8_t r[32], uint8_t out[32]) { uint8_t m[64]; memcpy(m, sk, 32); memcpy(m + 32, r, 32); sha256(m, 64, out); }
void compute_crh(const uint8_t pk[20], const uint8_t r[32], uint8_t out[32]) { uint8_t m[52]; memcpy(m, pk, 20); memcpy(m + 20, r, 32); sha256(m, 52, out); }

// full depth-8 tree with zero-padded leaves == IncrementalMerkleTree with EmptyMerkleRoots fillers (IncrementalMerkleTree.tcc)
static void merkle_levels(const uint8_t (*leaves)[32], size_t n, std::vector<std::vector<uint8_t>> &levels) {
    const size_t width = (size_t)1 << MERKLE_DEPTH;
    levels.assign(MERKLE_DEPTH + 1, std::vector<uint8_t>());
    levels[0].assign(width * 32, 0);
    for (size_t i = 0; i < n && i < width; i++) memcpy(&levels[0][i * 32], leaves[i], 32);
    for (int d = 1; d <= MERKLE_DEPTH; d++) {
        const size_t w = width >> d;
        levels[d].assign(w * 32, 0);
        for (size_t i = 0; i < w; i++) sha256_compress(&levels[d - 1][2 * i * 32], &levels[d - 1][(2 * i + 1) * 32], &levels[d][i * 32]);
    }
}
void merkle_root(const uint8_t (*leaves)[32], size_t n, uint8_t out[32]) {
    std::vector<std::vector<uint8_t>> lv; merkle_levels(leaves, n, lv); memcpy(out, lv[MERKLE_DEPTH].data(), 32);
}
void merkle_path(const uint8_t (*leaves)[32], size_t n, size_t index, uint8_t siblings[MERKLE_DEPTH][32], uint8_t root[32]) {
    std::vector<std::vector<uint8_t>> lv; merkle_levels(leaves, n, lv);
    size_t p = index;
    for (int d = 0; d < MERKLE_DEPTH; d++) { memcpy(siblings[d], &lv[d][(p ^ 1) * 32], 32); p >>= 1; }
    memcpy(root, lv[MERKLE_DEPTH].data(), 32);
}

// =====================================================================================================================
// the tape: variable values in allocation order
namespace {
typedef std::vector<int32_t> Refs;        // >= 0: variable index (0 = constant ONE); -1: the constant 0 (e.g. an IV bit that is 0)

struct Tape {
    std::vector<uint64_t> own;
    std::vector<WideValue> wide;
    uint64_t *v, *ext;
    uint32_t next = 1, cap;
    bool defer = false;                     // compressions record a seed instead of writing their 24 792 variables
    std::vector<CompressionSeed> seeds;
    Tape(uint32_t nvars, uint64_t *ext_, bool defer_ = false) : ext(ext_), cap(nvars), defer(defer_) {
        const size_t words = (size_t)nvars + 1;
        if (ext) { v = ext; if (!defer) memset(v, 0, words * 8); } else { own.assign(words, 0); v = own.data(); }
        v[0] = 1;
    }
    // deferred mode never touches the compression runs (97 % of the tape), so the tape is cleared piecewise as variables are handed out
    uint32_t alloc(uint32_t n = 1) { uint32_t r = next; next += n; if (defer && ext) memset(v + r, 0, (size_t)n * 8); return r; }
    uint32_t alloc_run(uint32_t n) { if (!defer) return alloc(n); uint32_t r = next; next += n; return r; }
    Refs alloc_refs(uint32_t n) { uint32_t b = alloc(n); Refs r(n); for (uint32_t i = 0; i < n; i++) r[i] = (int32_t)(b + i); return r; }
    uint64_t get(int32_t ref) const { return ref < 0 ? 0 : v[ref]; }
    void set(uint32_t idx, uint64_t x) { v[idx] = x; }
    void set4(uint32_t idx, const uint64_t x[4]) {                  // values that may exceed 64 bits (packed inputs, 2^64+B-A, inverses)
        v[idx] = x[0];
        for (size_t k = 0; k < wide.size(); k++) if (wide[k].idx == idx) { wide.erase(wide.begin() + k); break; }
        if (x[1] | x[2] | x[3]) { WideValue w; w.idx = idx; w.pad = 0; memcpy(w.v, x, 32); wide.push_back(w); }
    }
    void setref(int32_t ref, uint64_t x) { if (ref > 0) set((uint32_t)ref, x); }     // writes to ONE / constants are dropped
};

// uint256_to_bool_vector / uint160_to_bool_vector (circuit/utils.tcc:26-35): byte i of the blob, most significant bit first
static void fill_blob_bits(Tape &t, const Refs &bits, const uint8_t *blob) {
    for (size_t i = 0; i < bits.size(); i++) t.setref(bits[i], (blob[i >> 3] >> (7 - (i & 7))) & 1);
}
// uint64_to_bool_vector (utils.tcc:38-44): 8 little-endian bytes, each MSB first
static void fill_u64_bits(Tape &t, const Refs &bits, uint64_t v) {
    uint8_t b[8]; put_le64(b, v); fill_blob_bits(t, bits, b);
}
// get_field_element_from_bits_by_order (pb_variable.tcc:119-132) applied to those 64 bits == the integer itself
static uint64_t packed_u64_by_order(const Tape &t, const Refs &bits) {
    uint64_t r = 0;
    for (size_t i = 0; i < 8; i++) for (size_t j = 0; j < 8; j++) r = (r << 1) | (t.get(bits[64 - 1 - i * 8 - (7 - j)]) & 1);
    return r;
}
// multipacking_gadget::generate_r1cs_witness_from_bits with chunk = Fr::capacity() = 253 (basic_gadgets.tcc:100-106)
static void multipack(Tape &t, const Refs &bits, uint32_t packed_base) {
    const size_t chunk = 253, n = bits.size(), chunks = (n + chunk - 1) / chunk;
    for (size_t c = 0; c < chunks; c++) {
        uint64_t x[4] = {0, 0, 0, 0};
        for (size_t k = 0; k < chunk && c * chunk + k < n; k++)
            if (t.get(bits[c * chunk + k]) & 1) x[k >> 6] |= (uint64_t)1 << (k & 63);
        t.set4(packed_base + (uint32_t)c, x);
    }
}
static Refs concat(std::initializer_list<Refs> parts) { Refs r; for (auto &p : parts) r.insert(r.end(), p.begin(), p.end()); return r; }
static Refs slice(const Refs &a, size_t lo, size_t hi) { return Refs(a.begin() + lo, a.begin() + hi); }
// SHA-256 padding for a message of `msg_bits` bits whose last block already holds `used` bits: from_bits(..., ZERO) maps 1 -> ONE (ref 0),
// 0 -> the ZERO variable (commitment.tcc:32-58)
static Refs padding_refs(size_t msg_bits, size_t used_in_block, int32_t zero_var) {
    Refs r;
    const size_t total = 512 - used_in_block;
    for (size_t i = 0; i < total; i++) {
        int bit = 0;
        if (i == 0) bit = 1;
        else if (i >= total - 64) bit = (int)(((uint64_t)msg_bits >> (total - 1 - i)) & 1);
        r.push_back(bit ? 0 : zero_var);
    }
    return r;
}
static Refs iv_refs() {      // SHA256_default_IV (sha256_components.tcc:36-51): constants, MSB first per word
    Refs r(256);
    for (int i = 0; i < 256; i++) r[i] = ((IV256[i / 32] >> (31 - (i % 32))) & 1) ? 0 : -1;
    return r;
}

// sha256_compression_function_gadget (sha256_gadget.tcc:20-140, sha256_components.tcc, sha256_aux.tcc).
// Variables are allocated by its constructor in one contiguous run of 24792 starting at `base`:
//   [0,64) packed_W | 48 x 152 message-schedule blocks (i = 16..63) | 64 x 272 round blocks | 8 unreduced_output, 8 reduced_output, 8 overflow bits
struct Compression {
    static constexpr uint32_t VARS = COMPRESSION_VARS;
    uint32_t base;
    Refs prev, block, out;       // 256 / 512 / 256 refs in digest bit order
    Compression(Tape &t, const Refs &prev_output, const Refs &new_block, const Refs &output) : prev(prev_output), block(new_block), out(output) {
        base = t.alloc_run(VARS);
    }
    // 32 consecutive bit variables (least significant bit first) of one word
    static void set_bits(Tape &t, uint32_t at, uint64_t x, int n = 32) { uint64_t *o = t.v + at; for (int k = 0; k < n; k++) o[k] = (x >> k) & 1; }
    static uint32_t rotr(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }
    void witness(Tape &t) const {
        // Word-level evaluation: every bit variable of the gadget is a bit of one of the 32-bit words below, so the words are computed
        // natively and their bits stored in runs of 32 (the reference evaluates bit by bit through the protoboard).
        uint64_t W[64];                       // packed_W values
        for (int i = 0; i < 16; i++) {
            uint64_t x = 0;
            for (int k = 0; k < 32; k++) x |= (t.get(block[32 * i + 31 - k]) & 1) << k;
            W[i] = x;
        }
        if (t.defer) {
            // the run is filled on the GPU from this seed; only the digest bits (variables outside the run) are needed here
            CompressionSeed sd; sd.base = base;
            uint32_t st[8]; uint8_t blk[64];
            for (int i = 0; i < 16; i++) { sd.w[i] = (uint32_t)W[i]; blk[4 * i] = (uint8_t)(W[i] >> 24); blk[4 * i + 1] = (uint8_t)(W[i] >> 16); blk[4 * i + 2] = (uint8_t)(W[i] >> 8); blk[4 * i + 3] = (uint8_t)W[i]; }
            for (int r = 0; r < 8; r++) { uint32_t x = 0; for (int k = 0; k < 32; k++) x |= (uint32_t)(t.get(prev[32 * r + 31 - k]) & 1) << k; sd.h[r] = st[r] = x; }
            sha256_block(st, blk);
            // output slot o = i + 4s takes d[3-i] + new_a[63-i] (s = 0) or h[3-i] + new_e[63-i] (s = 1), i.e. digest word i + 4s (see the end of
            // the full evaluation below)
            for (int o = 0; o < 8; o++) for (int k = 0; k < 32; k++) t.setref(out[32 * o + 31 - k], (st[o] >> k) & 1);
            t.seeds.push_back(sd);
            return;
        }
        for (int i = 0; i < 16; i++) t.set(base + i, W[i]);
        for (int i = 16; i < 64; i++) {
            const uint32_t mb = base + 64 + (uint32_t)(i - 16) * 152;
            // small sigma0 on W[i-15] (7, 18, >>3), small sigma1 on W[i-2] (17, 19, >>10): 32 result bits, then the XOR3 tmps of the
            // positions that have a shifted-in third operand
            const uint32_t x0 = (uint32_t)W[i - 15], x1 = (uint32_t)W[i - 2];
            const uint32_t ab0 = rotr(x0, 7) ^ rotr(x0, 18), ab1 = rotr(x1, 17) ^ rotr(x1, 19);
            const uint64_t sig0 = ab0 ^ (x0 >> 3), sig1 = ab1 ^ (x1 >> 10);
            set_bits(t, mb + 2, sig0); set_bits(t, mb + 2 + 32, ab0, 32 - 3);
            const uint32_t c1 = mb + 2 + 32 + (32 - 3);
            set_bits(t, c1, sig1); set_bits(t, c1 + 32, ab1, 32 - 10);
            t.set(mb, sig0); t.set(mb + 1, sig1);
            const uint64_t unred = sig0 + sig1 + W[i - 16] + W[i - 7];
            t.set(mb + 117, unred);
            set_bits(t, mb + 118, unred, 34);       // 32 result bits, then the two overflow bits at mb + 150, mb + 151
            W[i] = unred & 0xffffffffull;
            t.set(base + i, W[i]);
        }
        // working variables
        uint32_t r8[8];
        for (int r = 0; r < 8; r++) { uint32_t x = 0; for (int k = 0; k < 32; k++) x |= (uint32_t)(t.get(prev[32 * r + 31 - k]) & 1) << k; r8[r] = x; }
        uint32_t a = r8[0], b = r8[1], c = r8[2], d = r8[3], e = r8[4], f = r8[5], g = r8[6], h = r8[7];
        const uint32_t rbase = base + 64 + 48 * 152;
        uint64_t packed_d[64], packed_h[64], packed_new_a[64], packed_new_e[64];
        for (int i = 0; i < 64; i++) {
            const uint32_t rb = rbase + (uint32_t)i * 272;
            // big sigma0 on a (2,13,22) and big sigma1 on e (6,11,25): 32 result bits then 32 XOR3 tmps each
            const uint32_t t0 = rotr(a, 2) ^ rotr(a, 13), t1 = rotr(e, 6) ^ rotr(e, 11);
            const uint64_t S0 = t0 ^ rotr(a, 22), S1 = t1 ^ rotr(e, 25);
            set_bits(t, rb + 66, S0); set_bits(t, rb + 66 + 32, t0);
            set_bits(t, rb + 130, S1); set_bits(t, rb + 130 + 32, t1);
            t.set(rb + 64, S0); t.set(rb + 65, S1);
            const uint64_t ch = (e & f) | (~e & g), mj = (a & b) | (a & c) | (b & c);
            set_bits(t, rb + 195, ch); set_bits(t, rb + 228, mj);
            t.set(rb + 194, ch); t.set(rb + 227, mj);
            packed_d[i] = d; packed_h[i] = h;
            t.set(rb + 260, packed_d[i]); t.set(rb + 261, packed_h[i]);
            const uint64_t ua = packed_h[i] + S1 + ch + K256[i] + W[i] + S0 + mj;
            const uint64_t ue = packed_d[i] + packed_h[i] + S1 + ch + K256[i] + W[i];
            t.set(rb + 262, ua); t.set(rb + 263, ue);
            set_bits(t, rb, ua); set_bits(t, rb + 32, ue);
            for (int k = 0; k < 3; k++) { t.set(rb + 266 + k, (ua >> (32 + k)) & 1); t.set(rb + 269 + k, (ue >> (32 + k)) & 1); }
            packed_new_a[i] = ua & 0xffffffffull; packed_new_e[i] = ue & 0xffffffffull;
            t.set(rb + 264, packed_new_a[i]); t.set(rb + 265, packed_new_e[i]);
            // rotate registers: h=g, g=f, f=e, e=new_e, d=c, c=b, b=a, a=new_a
            h = g; g = f; f = e; e = (uint32_t)packed_new_e[i]; d = c; c = b; b = a; a = (uint32_t)packed_new_a[i];
        }
        const uint32_t fb = rbase + 64 * 272;
        for (int i = 0; i < 4; i++) {
            const uint64_t u0 = packed_d[3 - i] + packed_new_a[63 - i], u1 = packed_h[3 - i] + packed_new_e[63 - i];
            const uint64_t us[2] = {u0, u1};
            for (int s = 0; s < 2; s++) {
                const int o = i + 4 * s;
                t.set(fb + o, us[s]);
                for (int k = 0; k < 32; k++) t.setref(out[32 * o + 31 - k], (us[s] >> k) & 1);
                t.set(fb + 16 + o, (us[s] >> 32) & 1);
                t.set(fb + 8 + o, us[s] & 0xffffffffull);
            }
        }
    }
};

// two-block hashers: intermediate digest (256 vars) then two compressions (commitment.tcc)
struct TwoBlockHash {
    Refs mid; Compression h1, h2;
    TwoBlockHash(Tape &t, Refs mid_, const Refs &block1, const Refs &block2, const Refs &out)
        : mid(std::move(mid_)), h1(t, iv_refs(), block1, mid), h2(t, mid, block2, out) {}
    void witness(Tape &t) const { h1.witness(t); h2.witness(t); }
};
static TwoBlockHash make_prf(Tape &t, int32_t zero, const Refs &sk, const Refs &rho, const Refs &out) {          // sha256_PRF_gadget
    Refs mid = t.alloc_refs(256);
    return TwoBlockHash(t, mid, concat({sk, rho}), padding_refs(512, 0, zero), out);
}
static TwoBlockHash make_cmta(Tape &t, int32_t zero, const Refs &v, const Refs &sn, const Refs &rho, const Refs &out) {   // sha256_CMTA_gadget
    Refs mid = t.alloc_refs(256);
    return TwoBlockHash(t, mid, concat({v, sn, slice(rho, 0, 192)}), concat({slice(rho, 192, 256), padding_refs(576, 64, zero)}), out);
}
static TwoBlockHash make_cmts(Tape &t, int32_t zero, const Refs &v, const Refs &pk, const Refs &r, const Refs &sn_old, const Refs &out) {   // sha256_CMTS_gadget
    Refs mid = t.alloc_refs(256);
    return TwoBlockHash(t, mid, concat({v, pk, r, slice(sn_old, 0, 32)}), concat({slice(sn_old, 32, 256), padding_refs(736, 224, zero)}), out);
}

// less_comparison_gadget (comparison.tcc:22-100): alpha[64], alpha_packed, not_all_zeros, disjunction inv
struct LessCmp {
    uint32_t alpha, alpha_packed, not_all_zeros, inv;
    explicit LessCmp(Tape &t) { alpha = t.alloc(64); alpha_packed = t.alloc(); not_all_zeros = t.alloc(); inv = t.alloc(); }
    void witness(Tape &t, uint64_t A, uint64_t B) const {       // packed(alpha) = 2^64 + B - A
        const unsigned __int128 x = ((unsigned __int128)1 << 64) + B - A;
        uint64_t limbs[4] = {(uint64_t)x, (uint64_t)(x >> 64), 0, 0};
        t.set4(alpha_packed, limbs);
        uint64_t pop = 0;
        for (int k = 0; k < 64; k++) { const uint64_t bit = (limbs[0] >> k) & 1; t.set(alpha + k, bit); pop += bit; }
        // alpha[64] is the constant ONE in the reference (alpha.emplace_back(0)): nothing to write.
        if (pop == 0) { t.set(inv, 0); t.set(not_all_zeros, 0); }
        else {
            uint64_t c[4]; zkh::HFr::from_u64(pop).inverse().to_canonical(c);
            t.set4(inv, c); t.set(not_all_zeros, 1);
        }
    }
};

static Assignment finish(Tape &t) {
    Assignment a; a.num_vars = t.cap; a.ext = t.ext; a.wide = std::move(t.wide); a.seeds = std::move(t.seeds); if (!t.ext) a.own = std::move(t.own); return a;
}
} // namespace

// =====================================================================================================================
// mint  (SRC/mint/circuit/gadget.tcc:71-162 allocation order; :194-246 witness order)
static Assignment mint_like(bool redeem, const Note &note_old, const Note &note, const uint8_t cmtA_old_d[32], const uint8_t cmtA_d[32],
                            uint64_t value_s_v, const uint8_t sk_d[32], uint64_t *ext, bool defer) {
    Tape t(redeem ? REDEEM_VARS : MINT_VARS, ext, defer);
    const uint32_t packed = t.alloc(4);
    Refs cmtA_old = t.alloc_refs(256), sn_old = t.alloc_refs(256), cmtA = t.alloc_refs(256), value_s = t.alloc_refs(64);
    const Refs unpacked = concat({cmtA_old, sn_old, cmtA, value_s});
    const uint32_t ZERO = t.alloc();
    Refs value = t.alloc_refs(64), value_old = t.alloc_refs(64);
    Refs sk = t.alloc_refs(256), r = t.alloc_refs(256), r_old = t.alloc_refs(256), sn = t.alloc_refs(256);
    const uint32_t value_packed = t.alloc(), value_old_packed = t.alloc(), value_s_packed = t.alloc();
    LessCmp *cmp = redeem ? new LessCmp(t) : nullptr;
    TwoBlockHash prf = make_prf(t, (int32_t)ZERO, sk, r, sn);
    TwoBlockHash cm_old = make_cmta(t, (int32_t)ZERO, value_old, sn_old, r_old, cmtA_old);
    TwoBlockHash cm = make_cmta(t, (int32_t)ZERO, value, sn, r, cmtA);

    // note_gadget_with_packing::generate_r1cs_witness (note.tcc:75-88)
    fill_u64_bits(t, value, note.value); t.set(value_packed, packed_u64_by_order(t, value));
    fill_u64_bits(t, value_old, note_old.value); t.set(value_old_packed, packed_u64_by_order(t, value_old));
    fill_u64_bits(t, value_s, value_s_v); t.set(value_s_packed, packed_u64_by_order(t, value_s));
    fill_blob_bits(t, sk, sk_d); fill_blob_bits(t, r, note.r); fill_blob_bits(t, r_old, note_old.r);
    if (redeem) {
        fill_blob_bits(t, sn, note.sn); fill_blob_bits(t, sn_old, note_old.sn);
        cmp->witness(t, t.get((int32_t)value_s_packed), t.get((int32_t)value_old_packed));
    }
    t.set(ZERO, 0);
    prf.witness(t);
    if (!redeem) { fill_blob_bits(t, sn, note.sn); fill_blob_bits(t, sn_old, note_old.sn); }
    cm_old.witness(t);
    cm.witness(t);
    fill_blob_bits(t, cmtA_old, cmtA_old_d);
    fill_blob_bits(t, cmtA, cmtA_d);
    multipack(t, unpacked, packed);
    delete cmp;
    return finish(t);
}
Assignment mint_witness(const Note &note_old, const Note &note, const uint8_t cmtA_old[32], const uint8_t cmtA[32], uint64_t value_s, const uint8_t sk[32], uint64_t *ext, bool defer) {
    return mint_like(false, note_old, note, cmtA_old, cmtA, value_s, sk, ext, defer);
}
Assignment redeem_witness(const Note &note_old, const Note &note, const uint8_t cmtA_old[32], const uint8_t cmtA[32], uint64_t value_s, const uint8_t sk[32], uint64_t *ext, bool defer) {
    return mint_like(true, note_old, note, cmtA_old, cmtA, value_s, sk, ext, defer);
}

// =====================================================================================================================
// send  (SRC/send/circuit/gadget.tcc constructor / generate_r1cs_witness; note.tcc; less_cmp.tcc; commitment.tcc)
Assignment send_witness(const Note &note_old, const NoteS &note_s, const Note &note, const uint8_t cmtA_old_d[32], const uint8_t cmtS_d[32],
                        const uint8_t cmtA_d[32], const uint8_t sk_d[32], const uint8_t pk_sender_d[20], uint64_t *ext, bool defer) {
    Tape t(SEND_VARS, ext, defer);
    const uint32_t packed = t.alloc(5);
    Refs cmtA_old = t.alloc_refs(256), sn_old = t.alloc_refs(256), cmtS = t.alloc_refs(256), cmtA = t.alloc_refs(256);
    const Refs unpacked = concat({cmtA_old, sn_old, cmtS, cmtA});
    const uint32_t ZERO = t.alloc();
    Refs value_old = t.alloc_refs(64), r_old = t.alloc_refs(256);
    Refs value_s = t.alloc_refs(64), pk_recv = t.alloc_refs(160), pk_sender = t.alloc_refs(160), r_s = t.alloc_refs(256);
    Refs value = t.alloc_refs(64), sn = t.alloc_refs(256), r = t.alloc_refs(256), sk = t.alloc_refs(256);
    // lessCMP: note_gadget_with_packing (value_old_packed, value_s_packed) + less_comparison_gadget
    const uint32_t c_vold_packed = t.alloc(), c_vs_packed = t.alloc();
    LessCmp cmp(t);
    // noteSUB: its own note_gadget_with_packing base + value_packed
    const uint32_t s_vold_packed = t.alloc(), s_vs_packed = t.alloc(), value_packed = t.alloc();
    // sha256_CRH_gadget: one block  pk_sender(160) | r(256) | padding(96), no intermediate digest
    Compression crh(t, iv_refs(), concat({pk_sender, r, padding_refs(416, 416, (int32_t)ZERO)}), r_s);
    TwoBlockHash prf = make_prf(t, (int32_t)ZERO, sk, r, sn);
    TwoBlockHash cm_old = make_cmta(t, (int32_t)ZERO, value_old, sn_old, r_old, cmtA_old);
    TwoBlockHash cm_s = make_cmts(t, (int32_t)ZERO, value_s, pk_recv, r_s, sn_old, cmtS);
    TwoBlockHash cm = make_cmta(t, (int32_t)ZERO, value, sn, r, cmtA);

    auto base_witness = [&](uint32_t vold_packed, uint32_t vs_packed) {       // note.tcc note_gadget_with_packing::generate_r1cs_witness
        fill_u64_bits(t, value_old, note_old.value); t.set(vold_packed, packed_u64_by_order(t, value_old));
        fill_blob_bits(t, sn_old, note_old.sn); fill_blob_bits(t, r_old, note_old.r);
        fill_u64_bits(t, value_s, note_s.value); t.set(vs_packed, packed_u64_by_order(t, value_s));
        fill_blob_bits(t, pk_recv, note_s.pk); fill_blob_bits(t, r_s, note_s.r);
    };
    base_witness(c_vold_packed, c_vs_packed);
    cmp.witness(t, t.get((int32_t)c_vs_packed), t.get((int32_t)c_vold_packed));
    base_witness(s_vold_packed, s_vs_packed);
    fill_u64_bits(t, value, note.value); t.set(value_packed, packed_u64_by_order(t, value));
    fill_blob_bits(t, sn, note.sn); fill_blob_bits(t, r, note.r); fill_blob_bits(t, sk, sk_d); fill_blob_bits(t, pk_sender, pk_sender_d);
    t.set(ZERO, 0);
    crh.witness(t);
    prf.witness(t);
    cm_old.witness(t); cm_s.witness(t); cm.witness(t);
    fill_blob_bits(t, cmtA_old, cmtA_old_d); fill_blob_bits(t, cmtS, cmtS_d); fill_blob_bits(t, cmtA, cmtA_d);
    multipack(t, unpacked, packed);
    return finish(t);
}

// =====================================================================================================================
// deposit  (SRC/deposit/circuit/gadget.tcc, merkle.tcc, note.tcc; libsnark merkle_tree_check_read_gadget.tcc:32-125)
Assignment deposit_witness(const NoteS &note_s, const Note &note_old, const Note &note, const uint8_t cmtS_d[32], const uint8_t cmtB_old_d[32],
                           const uint8_t cmtB_d[32], const uint8_t rt_d[32], size_t leaf_index, const uint8_t siblings[MERKLE_DEPTH][32],
                           const uint8_t sn_s_d[32], const uint8_t sk_d[32], uint64_t *ext, bool defer) {
    const int D = MERKLE_DEPTH;
    Tape t(DEPOSIT_VARS, ext, defer);
    const uint32_t packed = t.alloc(6);
    Refs root = t.alloc_refs(256), pk_recv = t.alloc_refs(160), cmtB_old = t.alloc_refs(256), sn_old = t.alloc_refs(256), cmtB = t.alloc_refs(256),
         sn_s = t.alloc_refs(256);
    const Refs unpacked = concat({root, pk_recv, cmtB_old, sn_old, cmtB, sn_s});
    const uint32_t value_enforce = t.alloc(), ZERO = t.alloc();
    Refs value_s = t.alloc_refs(64), r_s = t.alloc_refs(256), sn_A_old = t.alloc_refs(256), cmtS = t.alloc_refs(256);
    Refs value_old = t.alloc_refs(64), r_old = t.alloc_refs(256), value = t.alloc_refs(64), sn = t.alloc_refs(256), r = t.alloc_refs(256),
         sk = t.alloc_refs(256);
    const uint32_t value_s_packed = t.alloc(), value_old_packed = t.alloc(), value_packed = t.alloc();
    TwoBlockHash prf_sn = make_prf(t, (int32_t)ZERO, sk, r, sn);
    TwoBlockHash prf_sn_s = make_prf(t, (int32_t)ZERO, sk, r_s, sn_s);
    TwoBlockHash cm_s = make_cmts(t, (int32_t)ZERO, value_s, pk_recv, r_s, sn_A_old, cmtS);
    TwoBlockHash cm_old = make_cmta(t, (int32_t)ZERO, value_old, sn_old, r_old, cmtB_old);
    TwoBlockHash cm = make_cmta(t, (int32_t)ZERO, value, sn, r, cmtB);
    // merkle_tree_gadget: positions, authentication path variables, internal outputs, computed root, 8 hashers, check_root packers
    Refs positions = t.alloc_refs(D);
    std::vector<Refs> left(D), right(D);
    for (int i = 0; i < D; i++) { left[i] = t.alloc_refs(256); right[i] = t.alloc_refs(256); }
    std::vector<Refs> internal(D - 1);
    for (int i = 0; i < D - 1; i++) internal[i] = t.alloc_refs(256);
    Refs computed_root = t.alloc_refs(256);
    std::vector<Compression> hashers;
    hashers.reserve(D);
    for (int i = 0; i < D; i++) hashers.emplace_back(t, iv_refs(), concat({left[i], right[i]}), i == 0 ? computed_root : internal[i - 1]);
    const uint32_t packed_source = t.alloc(2), packed_target = t.alloc(2);

    // note_gadget_with_packing_and_ADD::generate_r1cs_witness (note.tcc)
    fill_u64_bits(t, value_s, note_s.value); t.set(value_s_packed, packed_u64_by_order(t, value_s));
    fill_u64_bits(t, value_old, note_old.value); t.set(value_old_packed, packed_u64_by_order(t, value_old));
    fill_u64_bits(t, value, note.value); t.set(value_packed, packed_u64_by_order(t, value));
    fill_blob_bits(t, pk_recv, note_s.pk); fill_blob_bits(t, r_s, note_s.r); fill_blob_bits(t, sn_A_old, note_s.sn_old);
    fill_blob_bits(t, sn_old, note_old.sn); fill_blob_bits(t, r_old, note_old.r);
    fill_blob_bits(t, sn, note.sn); fill_blob_bits(t, r, note.r); fill_blob_bits(t, sk, sk_d);
    t.set(value_enforce, note_s.value != 0 ? 1 : 0);
    t.set(ZERO, 0);
    prf_sn.witness(t);
    prf_sn_s.witness(t);
    fill_blob_bits(t, sn_s, sn_s_d);
    cm_s.witness(t); cm_old.witness(t); cm.witness(t);
    fill_blob_bits(t, cmtS, cmtS_d); fill_blob_bits(t, cmtB_old, cmtB_old_d); fill_blob_bits(t, cmtB, cmtB_d);
    // merkle_tree_gadget::generate_r1cs_witness (merkle.tcc): positions = bits of the leaf index, LSB first
    for (int i = 0; i < D; i++) t.set((uint32_t)positions[i], (leaf_index >> i) & 1);
    // merkle_authentication_path_variable::generate_r1cs_witness: level i counted from the ROOT; path[i] = sibling at tree level D-1-i
    for (int i = 0; i < D; i++) {
        const uint8_t *sib = siblings[D - 1 - i];
        if (leaf_index & ((size_t)1 << (D - 1 - i))) fill_blob_bits(t, left[i], sib); else fill_blob_bits(t, right[i], sib);
    }
    // merkle_tree_check_read_gadget::generate_r1cs_witness: bottom-up, propagate then hash
    for (int i = D - 1; i >= 0; i--) {
        const Refs &input = (i < D - 1) ? internal[i] : cmtS;
        const bool is_right = t.get(positions[D - 1 - i]) == 1;
        const Refs &dst = is_right ? right[i] : left[i];
        for (int k = 0; k < 256; k++) t.setref(dst[k], t.get(input[k]));
        hashers[i].witness(t);
    }
    // bit_vector_copy_gadget (basic_gadgets.tcc:180-194): copy when value_enforce == 1, then pack source and target
    if (t.get((int32_t)value_enforce) == 1) for (int k = 0; k < 256; k++) t.setref(root[k], t.get(computed_root[k]));
    multipack(t, computed_root, packed_source);
    multipack(t, root, packed_target);
    fill_blob_bits(t, root, rt_d);
    multipack(t, unpacked, packed);
    return finish(t);
}

} // namespace zkw
