// C-ABI layer (1) of include/zkb200.h: BlockMaze's cgo surface (libzk_mint / libzk_send / libzk_deposit / libzk_redeem),
// re-exported from one library.  Argument parsing, returned-buffer semantics and the failure encoding follow
// SRC/{mint,send,deposit,redeem}/*cgo.cpp; the proving itself runs on the GPU against proving keys that are parsed once per
// process and stay resident (the reference re-reads and re-parses the key file inside every gen*proof call,
// mintcgo.cpp:299-302).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include "../../include/zkb200.h"
#include "prover.cuh"
#include "sched.hpp"
#include "witness.hpp"

using namespace zkw;

// Everything the cgo layer keeps between calls.  One geth process drives EVERY GPU of the box: a
// circuit's proving key is parsed once and made resident on each active device, and every gen*proof call goes to the device with the
// fewest proofs in flight (zkb::DeviceSched), then to a free lane of that device's copy of the key.  The reference's only parallelism
// knob is the OpenMP chunk count of its multi_exp (r1cs_gg_ppzksnark.tcc:429-433); zktx.go:383-404 stays unchanged.
struct KeySet {                                  // one circuit, one resident copy per active device
    std::vector<int> devices;
    std::vector<void *> pk;
    ~KeySet() { for (void *p : pk) if (p) zkb200_pk_free(p); }
};
static std::mutex g_abi_mu;                      // key directory, device list, random stream
static std::mutex g_key_mu[4];                   // loading one circuit's keys does not block the callers of another
static std::string g_key_dir;
static std::vector<int> g_devices;               // active devices of the cgo layer (empty = not decided yet)
static std::shared_ptr<zkb::DeviceSched> g_sched;
static std::shared_ptr<KeySet> g_keys[4];
static std::vector<uint32_t> g_words;
static size_t g_word_pos = 0;
static thread_local double g_last_ms[4] = {0, 0, 0, 0};      // witness generation, prove call, of which GPU, host finish (last gen*proof of this thread)
static thread_local int g_last_device = -1, g_last_lane = -1;
static std::atomic<long> g_device_proofs[64];                                      // gen*proof calls finished per device (zkb200_device_proofs)
static thread_local void *g_last_pk = nullptr;
// SHA-256 gadget runs expanded on the GPU (default) or by the host generator (ZKB200_GPU_WITNESS=0)
static bool gpu_witness() { static const bool on = [] { const char *e = getenv("ZKB200_GPU_WITNESS"); return !(e && atoi(e) == 0); }(); return on; }
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static const char *CIRCUIT_NAMES[4] = {"mint", "send", "deposit", "redeem"};

std::string zkw::key_dir() {
    std::lock_guard<std::mutex> lk(g_abi_mu);
    if (!g_key_dir.empty()) return g_key_dir;
    const char *e = getenv("ZKB200_KEY_DIR");
    return e ? std::string(e) : std::string("/usr/local/prfKey");        // hard-coded in the reference (mintcgo.cpp:302)
}

static void drop_keys() {                        // resident keys are reference-counted: a proof in flight keeps its KeySet alive
    for (int c = 0; c < 4; c++) { std::lock_guard<std::mutex> lk(g_key_mu[c]); g_keys[c].reset(); }
}
void zkb200_set_key_dir(const char *dir) {
    { std::lock_guard<std::mutex> lk(g_abi_mu); g_key_dir = dir ? dir : ""; }
    drop_keys();
}
// The devices gen*proof spreads over: zkb200_set_devices() > $ZKB200_DEVICES ("all" or "0,2,3") > the device of an explicit zkb200_init()
// (one process per GPU, as bench.py under torchrun) > $ZKB200_DEVICE > every visible device.
static std::vector<int> decide_devices() {
    std::vector<int> out;
    const int count = zkb200_device_count();
    if (const char *e = getenv("ZKB200_DEVICES")) {
        if (strcmp(e, "all") != 0) {
            for (const char *p = e; *p;) { char *q; const long d = strtol(p, &q, 10); if (q == p) break; if (d >= 0 && d < count) out.push_back((int)d); p = (*q == ',') ? q + 1 : q; }
        }
        if (out.empty()) for (int d = 0; d < count; d++) out.push_back(d);
        return out;
    }
    const int chosen = zkb200_current_device();
    if (chosen >= 0) return {chosen};
    if (const char *e = getenv("ZKB200_DEVICE")) { const int d = atoi(e); if (d >= 0 && d < count) return {d}; }
    for (int d = 0; d < count; d++) out.push_back(d);
    return out;
}
int zkb200_set_devices(const int *devices, int n) {
    const int count = zkb200_device_count();
    std::vector<int> v;
    for (int i = 0; i < n; i++) { if (devices[i] < 0 || devices[i] >= count) return -1; v.push_back(devices[i]); }
    if (n == 0) for (int d = 0; d < count; d++) v.push_back(d);
    if (v.empty()) return -1;
    { std::lock_guard<std::mutex> lk(g_abi_mu); g_devices = v; g_sched = std::make_shared<zkb::DeviceSched>((int)v.size()); }
    drop_keys();
    return (int)v.size();
}
int zkb200_active_devices(int *out, int cap) {
    std::lock_guard<std::mutex> lk(g_abi_mu);
    if (g_devices.empty()) { g_devices = decide_devices(); g_sched = std::make_shared<zkb::DeviceSched>((int)g_devices.size()); }
    for (int i = 0; i < (int)g_devices.size() && i < cap; i++) out[i] = g_devices[i];
    return (int)g_devices.size();
}

// Test hook: pins the prover randomness.  It only exists when the process was started with ZKB200_TEST_RNG=1 -- a pinned (r, s) reused
// across proofs leaks witness relations, so production processes cannot switch it on by accident -- and it never wraps: once the words
// are used up the prover is back on std::random_device.
int zkb200_set_random_words(const uint32_t *words, size_t n_words) {
    const char *e = getenv("ZKB200_TEST_RNG");
    if (n_words && !(e && atoi(e) == 1)) {
        fprintf(stderr, "zkb200: zkb200_set_random_words refused (set ZKB200_TEST_RNG=1 in test processes only)\n");
        return -1;
    }
    std::lock_guard<std::mutex> lk(g_abi_mu);
    g_words.assign(words, words + n_words);
    g_word_pos = 0;
    return 0;
}

// Fr::random_element (fp.tcc:695-721, bigint.tcc:167-179): 8 x 32-bit words -> mont_repr, clear bits >= 254, retry while >= r.
// The field element is the one whose MONTGOMERY representative is that integer; returns its canonical value.  Caller holds g_abi_mu.
static void next_random_fr(uint64_t out[4]) {
    static std::random_device rd;
    for (;;) {
        uint32_t w[8];
        for (int i = 0; i < 8; i++) {
            if (g_word_pos < g_words.size()) w[i] = g_words[g_word_pos++];
            else { if (!g_words.empty()) { g_words.clear(); g_word_pos = 0; } w[i] = rd(); }
        }
        uint64_t m[4];
        for (int i = 0; i < 4; i++) m[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
        m[3] &= 0x3fffffffffffffffull;
        if (!zkh::HFr::geq_mod(m)) { zkh::HFr::raw(m).to_canonical(out); return; }
    }
}

// the circuit's keys on every active device; parsed once, uploaded by one thread per device
static std::shared_ptr<KeySet> circuit_keys(int circuit, std::shared_ptr<zkb::DeviceSched> &sched) {
    int devs[64];
    const int nd = zkb200_active_devices(devs, 64);
    { std::lock_guard<std::mutex> lk(g_abi_mu); sched = g_sched; }
    std::lock_guard<std::mutex> lk(g_key_mu[circuit]);
    if (g_keys[circuit] && (int)g_keys[circuit]->devices.size() == nd && std::equal(devs, devs + nd, g_keys[circuit]->devices.begin())) return g_keys[circuit];
    const std::string path = key_dir() + "/" + CIRCUIT_NAMES[circuit] + "pk.txt";
    auto ks = std::make_shared<KeySet>();
    ks->devices.assign(devs, devs + nd);
    ks->pk.assign(nd, nullptr);
    if (zkb200_pk_load_many(path.c_str(), devs, nd, ks->pk.data()) != 0) {
        // the reference's behaviour on a missing key file is undefined (assert compiled out, mintcgo.cpp:69); fail loudly instead
        fprintf(stderr, "zkb200: cannot load %s: %s\n", path.c_str(), zkb200_last_error());
        abort();
    }
    g_keys[circuit] = ks;
    return ks;
}

static char *dup_hex(const std::string &s, size_t cap) {       // `new char[cap]`, NUL-terminated, like the reference helpers
    char *p = new char[cap];
    memset(p, 0, cap);
    memcpy(p, s.data(), s.size() < cap - 1 ? s.size() : cap - 1);
    return p;
}
// gen*proof core: pick the least-loaded device, take a free lane of its copy of the key, generate the witness straight into that lane's
// pinned staging buffer (no intermediate copy), prove on it.  Only the key lookup and the random draws are serialised: concurrent callers
// overlap, on one GPU and across GPUs.
template <class Fn> static char *prove_timed(int circuit, Fn make) {
    std::shared_ptr<zkb::DeviceSched> sched;
    const std::shared_ptr<KeySet> ks = circuit_keys(circuit, sched);          // held for the whole call: zkb200_set_key_dir cannot free it under us
    const int slot = sched->pick();
    void *pk = ks->pk[slot];
    g_last_device = ks->devices[slot]; g_last_pk = pk;
    const int lane = zkb200_lane_acquire(pk);
    uint64_t *ext = zkb200_lane_staging(pk, lane);             // pinned: the generator writes the compact assignment in place
    const double t0 = now_ms();
    const Assignment a = make(ext);
    g_last_ms[0] = now_ms() - t0;
    uint64_t r[4], s[4];
    { std::lock_guard<std::mutex> lk(g_abi_mu); next_random_fr(r); next_random_fr(s); }      // r first, then s (r1cs_gg_ppzksnark.tcc:418-419)
    char *p = new char[1153];                                  // the reference returns new char[1153] with 512 hex chars (mintcgo.cpp:316-320)
    memset(p, 0, 1153);
    float tm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double t1 = now_ms();
    const int rc = zkb200_prove_compact(pk, a.lo(), a.wide.data(), a.wide.size(), a.seeds.data(), a.seeds.size(), (const uint8_t *)r, (const uint8_t *)s, p, tm);
    g_last_lane = lane;
    g_last_ms[1] = now_ms() - t1; g_last_ms[2] = tm[0]; g_last_ms[3] = tm[3];
    zkb200_lane_release(pk, lane);
    sched->done(slot);
    if (g_last_device >= 0 && g_last_device < 64) g_device_proofs[g_last_device].fetch_add(1, std::memory_order_relaxed);
    if (rc < 0) { fprintf(stderr, "zkb200: prover rejected the %s witness (%d)\n", CIRCUIT_NAMES[circuit], rc); abort(); }   // generator bug, never a wrong proof
    if (rc == 1) printf("can not generate %s proof\n", CIRCUIT_NAMES[circuit]);      // mintcgo.cpp:209
    return p;
}
int zkb200_last_device(void) { return g_last_device; }
long zkb200_device_proofs(int device) { return device >= 0 && device < 64 ? g_device_proofs[device].load(std::memory_order_relaxed) : -1; }
// parity hook: the assignment on the GPU behind this thread's last gen*proof (valid until another proof takes that lane)
long zkb200_last_assignment(uint8_t *out, size_t cap_bytes) {
    if (!g_last_pk || g_last_lane < 0) return -1;
    uint64_t info[8];
    if (zkb200_pk_info(g_last_pk, info, nullptr)) return -1;
    if (cap_bytes < info[0] * 32) return -1;
    return zkb200_lane_read_assignment(g_last_pk, g_last_lane, out) ? -1 : (long)info[0];
}

// ---- helpers ---------------------------------------------------------------------------------------------------------------
char *genCMT(uint64_t value, char *sn_string, char *r_string) {
    uint8_t sn[32], r[32], cm[32];
    parse_hex_blob(sn_string, sn, 32); parse_hex_blob(r_string, r, 32);
    note_cm(value, sn, r, cm);
    return dup_hex(blob_to_hex(cm, 32), 68);
}
char *computePRF(char *sk_string, char *r_string) {
    uint8_t sk[32], r[32], o[32];
    parse_hex_blob(sk_string, sk, 32); parse_hex_blob(r_string, r, 32);
    compute_prf(sk, r, o);
    return dup_hex(blob_to_hex(o, 32), 68);
}
char *genCMTS(uint64_t value_s, char *pk_string, char *r_s_string, char *sn_old_string) {
    uint8_t pk[20], r[32], sn[32], cm[32];
    parse_hex_blob(pk_string, pk, 20); parse_hex_blob(r_s_string, r, 32); parse_hex_blob(sn_old_string, sn, 32);
    notes_cm(value_s, pk, r, sn, cm);
    return dup_hex(blob_to_hex(cm, 32), 68);
}
char *computeCRH(char *pk_string, char *r_string) {
    uint8_t pk[20], r[32], o[32];
    parse_hex_blob(pk_string, pk, 20); parse_hex_blob(r_string, r, 32);
    compute_crh(pk, r, o);
    return dup_hex(blob_to_hex(o, 32), 68);
}
static size_t parse_cmtarray(const char *cmtarray, int n, std::vector<uint8_t> &leaves) {
    // n concatenated 66-character strings "0x" + 64 hex (depositcgo.cpp:304-315); the reference holds at most 256 of them
    const size_t cnt = n < 0 ? 0 : (n > 256 ? 256 : (size_t)n);
    leaves.assign(cnt * 32, 0);
    const size_t len = cmtarray ? strlen(cmtarray) : 0;
    for (size_t i = 0; i < cnt; i++) {
        if (i * 66 >= len) break;
        std::string piece(cmtarray + i * 66, std::min<size_t>(66, len - i * 66));
        parse_hex_blob(piece.c_str(), &leaves[i * 32], 32);
    }
    return cnt;
}
char *genRoot(char *cmtarray, int n) {
    std::vector<uint8_t> leaves;
    const size_t cnt = parse_cmtarray(cmtarray, n, leaves);
    uint8_t rt[32];
    merkle_root((const uint8_t(*)[32])leaves.data(), cnt, rt);
    return dup_hex(blob_to_hex(rt, 32), 68);
}

// ---- proofs ------------------------------------------------------------------------------------------------------------------
static void parse_note(Note &n, uint64_t value, const char *sn, const char *r) { n.value = value; parse_hex_blob(sn, n.sn, 32); parse_hex_blob(r, n.r, 32); }

char *genMintproof(uint64_t value, uint64_t value_old, char *sn_old_string, char *r_old_string, char *sn_string, char *r_string,
                   char *cmtA_old_string, char *cmtA_string, uint64_t value_s, char *sk_string) {
    Note note_old, note; uint8_t cmtA_old[32], cmtA[32], sk[32];
    parse_note(note_old, value_old, sn_old_string, r_old_string); parse_note(note, value, sn_string, r_string);
    parse_hex_blob(cmtA_old_string, cmtA_old, 32); parse_hex_blob(cmtA_string, cmtA, 32); parse_hex_blob(sk_string, sk, 32);
    return prove_timed(ZKB200_MINT, [&](uint64_t *ext) { return mint_witness(note_old, note, cmtA_old, cmtA, value_s, sk, ext, gpu_witness()); });
}
char *genRedeemproof(uint64_t value, uint64_t value_old, char *sn_old_string, char *r_old_string, char *sn_string, char *r_string,
                     char *cmtA_old_string, char *cmtA_string, uint64_t value_s, char *sk_string) {
    Note note_old, note; uint8_t cmtA_old[32], cmtA[32], sk[32];
    parse_note(note_old, value_old, sn_old_string, r_old_string); parse_note(note, value, sn_string, r_string);
    parse_hex_blob(cmtA_old_string, cmtA_old, 32); parse_hex_blob(cmtA_string, cmtA, 32); parse_hex_blob(sk_string, sk, 32);
    printf("Trying to generate redeem proof...\n");            // redeemcgo.cpp:310
    return prove_timed(ZKB200_REDEEM, [&](uint64_t *ext) { return redeem_witness(note_old, note, cmtA_old, cmtA, value_s, sk, ext, gpu_witness()); });
}
char *genSendproof(uint64_t value_A, char *r_s_string, char *sn_string, char *r_string, char *cmt_s_string, char *cmtA_string, uint64_t value_s,
                   char *pk_recv_string, uint64_t value_A_new, char *sn_A_new, char *r_A_new, char *cmt_A_new, char *sk_string,
                   char *pk_sender_string) {
    // sendcgo.cpp:318-334: the "A" note is the old one; NoteS(value_s, pk_recv, r_s, sn)
    Note note_old, note_new; NoteS notes; uint8_t cmtS[32], cmtA[32], cmtAnew[32], sk[32], pk_sender[20];
    parse_note(note_old, value_A, sn_string, r_string); parse_note(note_new, value_A_new, sn_A_new, r_A_new);
    notes.value = value_s; parse_hex_blob(pk_recv_string, notes.pk, 20); parse_hex_blob(r_s_string, notes.r, 32); memcpy(notes.sn_old, note_old.sn, 32);
    parse_hex_blob(cmt_s_string, cmtS, 32); parse_hex_blob(cmtA_string, cmtA, 32); parse_hex_blob(cmt_A_new, cmtAnew, 32);
    parse_hex_blob(sk_string, sk, 32); parse_hex_blob(pk_sender_string, pk_sender, 20);
    printf("Trying to generate send proof...\n");              // sendcgo.cpp:352
    return prove_timed(ZKB200_SEND, [&](uint64_t *ext) { return send_witness(note_old, notes, note_new, cmtA, cmtS, cmtAnew, sk, pk_sender, ext, gpu_witness()); });
}
char *genDepositproof(uint64_t value, uint64_t value_old, char *sn_old_string, char *r_old_string, char *sn_string, char *r_string,
                      char *sns_string, char *rs_string, char *cmtB_old_string, char *cmtB_string, uint64_t value_s, char *pk_string,
                      char *sn_A_oldstring, char *cmtS_string, char *cmtarray, int n, char *RT, char *sk_string) {
    (void)RT;                                                  // accepted but ignored by the reference too (depositcgo.cpp:402-403)
    Note note_old, note; NoteS note_s; uint8_t sn_s[32], cmtB_old[32], cmtB[32], cmtS[32], sk[32];
    parse_note(note_old, value_old, sn_old_string, r_old_string); parse_note(note, value, sn_string, r_string);
    note_s.value = value_s; parse_hex_blob(pk_string, note_s.pk, 20); parse_hex_blob(rs_string, note_s.r, 32);
    parse_hex_blob(sn_A_oldstring, note_s.sn_old, 32);
    parse_hex_blob(sns_string, sn_s, 32); parse_hex_blob(cmtB_old_string, cmtB_old, 32); parse_hex_blob(cmtB_string, cmtB, 32);
    parse_hex_blob(cmtS_string, cmtS, 32); parse_hex_blob(sk_string, sk, 32);
    std::vector<uint8_t> leaves;
    const size_t cnt = parse_cmtarray(cmtarray, n, leaves);
    // depositcgo.cpp:373-400: leaves up to the FIRST occurrence of cmtS go to the tree; the witness is re-created at every later
    // occurrence (dropping what it had collected since), so the effective leaf list is  leaves[0..first] ++ leaves[last+1..]
    long first = -1, last = -1;
    for (size_t i = 0; i < cnt; i++) if (memcmp(&leaves[i * 32], cmtS, 32) == 0) { if (first < 0) first = (long)i; last = (long)i; }
    printf("Trying to generate deposit proof...\n");           // depositcgo.cpp:421
    if (first < 0) {
        // the reference throws from IncrementalMerkleTree::path() here (uncaught C++ exception under cgo); report failure instead
        printf("can not generate deposit proof\n");
        char *p = new char[1153]; memset(p, 0, 1153);
        memcpy(p, zkb200_default_proof(), 512);                 // r1cs_gg_ppzksnark_proof default constructor: the three generators
        return p;
    }
    std::vector<uint8_t> eff(leaves.begin(), leaves.begin() + (first + 1) * 32);
    eff.insert(eff.end(), leaves.begin() + (last + 1) * 32, leaves.end());
    uint8_t siblings[MERKLE_DEPTH][32], rt[32];
    merkle_path((const uint8_t(*)[32])eff.data(), eff.size() / 32, (size_t)first, siblings, rt);
    return prove_timed(ZKB200_DEPOSIT, [&](uint64_t *ext) { return deposit_witness(note_s, note_old, note, cmtS, cmtB_old, cmtB, rt, (size_t)first, siblings, sn_s, sk, ext, gpu_witness()); });
}

void zkb200_last_breakdown_ms(double out[4]) { for (int i = 0; i < 4; i++) out[i] = g_last_ms[i]; }

// parity hook for the GPU witness path: when on, the zkb200_witness_* functions below run the generator in its deferred mode (SHA-256 runs
// left out, seeds recorded) and fill the runs in with the host copy of the expansion the GPU kernel performs (witness_sha.hpp)
static thread_local bool g_witness_defer = false;
void zkb200_witness_defer(int on) { g_witness_defer = on != 0; }

// ---- witness-only entry points (parity hooks for the assignment layout; declared extern "C" in zkb200.h) ------------------------
long zkb200_witness_mint(uint64_t value, uint64_t value_old, const char *sn_old, const char *r_old, const char *sn, const char *r,
                         const char *cmtA_old_s, const char *cmtA_s, uint64_t value_s, const char *sk_s, int redeem, uint8_t *out, size_t cap) {
    Note note_old, note; uint8_t cmtA_old[32], cmtA[32], sk[32];
    parse_note(note_old, value_old, sn_old, r_old); parse_note(note, value, sn, r);
    parse_hex_blob(cmtA_old_s, cmtA_old, 32); parse_hex_blob(cmtA_s, cmtA, 32); parse_hex_blob(sk_s, sk, 32);
    Assignment a = redeem ? redeem_witness(note_old, note, cmtA_old, cmtA, value_s, sk, nullptr, g_witness_defer) : mint_witness(note_old, note, cmtA_old, cmtA, value_s, sk, nullptr, g_witness_defer);
    if (cap < a.num_vars) return -1;
    a.expand(out);
    return a.num_vars;
}
long zkb200_witness_send(uint64_t value_A, const char *r_s, const char *sn, const char *r, const char *cmt_s, const char *cmtA_s, uint64_t value_s,
                         const char *pk_recv, uint64_t value_A_new, const char *sn_A_new, const char *r_A_new, const char *cmt_A_new, const char *sk_s,
                         const char *pk_sender_s, uint8_t *out, size_t cap) {
    Note note_old, note_new; NoteS notes; uint8_t cmtS[32], cmtA[32], cmtAnew[32], sk[32], pk_sender[20];
    parse_note(note_old, value_A, sn, r); parse_note(note_new, value_A_new, sn_A_new, r_A_new);
    notes.value = value_s; parse_hex_blob(pk_recv, notes.pk, 20); parse_hex_blob(r_s, notes.r, 32); memcpy(notes.sn_old, note_old.sn, 32);
    parse_hex_blob(cmt_s, cmtS, 32); parse_hex_blob(cmtA_s, cmtA, 32); parse_hex_blob(cmt_A_new, cmtAnew, 32);
    parse_hex_blob(sk_s, sk, 32); parse_hex_blob(pk_sender_s, pk_sender, 20);
    Assignment a = send_witness(note_old, notes, note_new, cmtA, cmtS, cmtAnew, sk, pk_sender, nullptr, g_witness_defer);
    if (cap < a.num_vars) return -1;
    a.expand(out);
    return a.num_vars;
}
long zkb200_witness_deposit(uint64_t value, uint64_t value_old, const char *sn_old, const char *r_old, const char *sn, const char *r, const char *sns,
                            const char *rs, const char *cmtB_old_s, const char *cmtB_s, uint64_t value_s, const char *pk, const char *sn_A_old,
                            const char *cmtS_s, const char *cmtarray, int n, const char *RT, const char *sk_s, uint8_t *out, size_t cap) {
    (void)RT;
    Note note_old, note; NoteS note_s; uint8_t sn_s[32], cmtB_old[32], cmtB[32], cmtS[32], sk[32];
    parse_note(note_old, value_old, sn_old, r_old); parse_note(note, value, sn, r);
    note_s.value = value_s; parse_hex_blob(pk, note_s.pk, 20); parse_hex_blob(rs, note_s.r, 32); parse_hex_blob(sn_A_old, note_s.sn_old, 32);
    parse_hex_blob(sns, sn_s, 32); parse_hex_blob(cmtB_old_s, cmtB_old, 32); parse_hex_blob(cmtB_s, cmtB, 32);
    parse_hex_blob(cmtS_s, cmtS, 32); parse_hex_blob(sk_s, sk, 32);
    std::vector<uint8_t> leaves;
    const size_t cnt = parse_cmtarray(cmtarray, n, leaves);
    long first = -1, last = -1;
    for (size_t i = 0; i < cnt; i++) if (memcmp(&leaves[i * 32], cmtS, 32) == 0) { if (first < 0) first = (long)i; last = (long)i; }
    if (first < 0) return -2;
    std::vector<uint8_t> eff(leaves.begin(), leaves.begin() + (first + 1) * 32);
    eff.insert(eff.end(), leaves.begin() + (last + 1) * 32, leaves.end());
    uint8_t siblings[MERKLE_DEPTH][32], rt[32];
    merkle_path((const uint8_t(*)[32])eff.data(), eff.size() / 32, (size_t)first, siblings, rt);
    Assignment a = deposit_witness(note_s, note_old, note, cmtS, cmtB_old, cmtB, rt, (size_t)first, siblings, sn_s, sk, nullptr, g_witness_defer);
    if (cap < a.num_vars) return -1;
    a.expand(out);
    return a.num_vars;
}

// ---- verification: see verifier.cpp -------------------------------------------------------------------------------------------

// ---- batch entry point: many transactions, every GPU ----------------------------------------------------------------------------
// A pool of caller threads (default: active devices x lanes of a key) drains the list through the same gen*proof functions a single
// geth goroutine calls; the device scheduler spreads them, the lanes overlap them on each GPU.
int zkb200_prove_batch(size_t n, const zkb200_tx *txs, char *proofs, int threads) {
    if (n && (!txs || !proofs)) return -1;
    for (size_t i = 0; i < n; i++) if (txs[i].circuit < 0 || txs[i].circuit > 3) return -1;
    if (threads <= 0) {
        int devs[64];
        const int nd = zkb200_active_devices(devs, 64);
        int lanes = 3;
        if (const char *e = getenv("ZKB200_LANES")) lanes = atoi(e) > 0 ? atoi(e) : 3;
        threads = nd * lanes;
    }
    if ((size_t)threads > n) threads = (int)n;
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n) return;
            const zkb200_tx &t = txs[i];
            char **s = (char **)t.s;
            char *p = nullptr;
            switch (t.circuit) {
            case ZKB200_MINT: p = genMintproof(t.u[0], t.u[1], s[0], s[1], s[2], s[3], s[4], s[5], t.u[2], s[6]); break;
            case ZKB200_REDEEM: p = genRedeemproof(t.u[0], t.u[1], s[0], s[1], s[2], s[3], s[4], s[5], t.u[2], s[6]); break;
            case ZKB200_SEND: p = genSendproof(t.u[0], s[0], s[1], s[2], s[3], s[4], t.u[1], s[5], t.u[2], s[6], s[7], s[8], s[9], s[10]); break;
            default: p = genDepositproof(t.u[0], t.u[1], s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], t.u[2], s[8], s[9], s[10], s[11], t.n, s[12], s[13]); break;
            }
            memcpy(proofs + 513 * i, p, 512); proofs[513 * i + 512] = 0;
            if (memcmp(p, "0000000000", 10) == 0) failed++;
            delete[] p;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    return failed.load();
}

// ---- device scheduler hooks (CPU unit tests of the dispatch policy) ---------------------------------------------------------------
void *zkb200_sched_create(int n_devices) { return new zkb::DeviceSched(n_devices); }
int zkb200_sched_pick(void *h) { return ((zkb::DeviceSched *)h)->pick(); }
void zkb200_sched_done(void *h, int slot) { ((zkb::DeviceSched *)h)->done(slot); }
int zkb200_sched_inflight(void *h, int slot) { return ((zkb::DeviceSched *)h)->inflight(slot); }
void zkb200_sched_free(void *h) { delete (zkb::DeviceSched *)h; }
