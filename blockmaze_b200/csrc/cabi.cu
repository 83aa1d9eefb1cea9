// C-ABI layer (2) of include/zkb200.h: resident proving keys, prover, kernel-level entry points and device benchmarks.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/zkb200.h"
#include "prover.cuh"
#include "ntt.cuh"
#include "msm.cuh"

using namespace zk;
using namespace zkp;

static std::string g_err;
static std::mutex g_err_mu;
static void set_err(const std::string &e) { std::lock_guard<std::mutex> lk(g_err_mu); g_err = e; }
// Device of the layer-2 calls that take no key handle (kernel entry points, benches).  Everything that takes a key handle runs on the
// device that key is resident on, and the cgo layer spreads over all active devices (blockmaze_abi.cu) -- nothing here pins a process to one GPU.
static int g_device = -1;
static bool g_explicit = false;              // zkb200_init() was called by the user (not by ensure_device)
static std::mutex g_mu;

static int init_device(int device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { set_err("no CUDA device: the B200 prover has no CPU fallback"); return -1; }
    if (device < 0 || device >= count) { set_err("bad device index"); return -1; }
    g_device = device;
    device_init(device);
    ntt_init_attrs();            // this translation unit's own instances of the NTT kernels (function attributes are per device)
    return 0;
}
static int ensure_device() {
    if (g_device >= 0) { cudaSetDevice(g_device); return 0; }
    const char *e = getenv("ZKB200_DEVICE");
    return init_device(e ? atoi(e) : 0);
}

template <class F> __global__ void to_mont_generic_kernel(F *a, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = a[i].to_mont();
}
template <class F> __global__ void field_op_kernel(const F *a, const F *b, F *out, size_t n, int op) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b ? b[i] : F::zero(), z;
    switch (op) {
    case 0: z = x * y; break;
    case 1: z = x + y; break;
    case 2: z = x - y; break;
    case 3: z = x.sqr(); break;
    case 4: z = x.to_mont(); break;
    case 5: z = x.from_mont(); break;
    default: z = x.inverse(); break;
    }
    out[i] = z;
}

// ---- synthetic inputs of the kernel sweep (BASELINE.json configs[4]) -------------------------------------------------
// Scalars are libff's SHA512_rng<Fr>(idx) (libff/common/rng.tcc:26-72): SHA-512 of (idx:u64 LE || iter:u64 LE), the first 32 digest bytes
// read as a little-endian integer, bits >= 254 cleared, next iter while the value is not below r.  Bases are P_i = h_i * G with
// h_i = SHA512_rng(2^32 + i): distinct points with no structure a bucket method could profit from (the reference's own profile repeats one
// point, multiexp_profile.cpp:20-31).  Both are functions of the GLOBAL index, so a slice of the problem can be made on any GPU, and the
// host side of the sweep feeds exactly these bytes to libff / libfqfft.
__device__ __constant__ uint64_t SHA512_K[80] = {
    0x428a2f98d728ae22ull, 0x7137449123ef65cdull, 0xb5c0fbcfec4d3b2full, 0xe9b5dba58189dbbcull, 0x3956c25bf348b538ull, 0x59f111f1b605d019ull, 0x923f82a4af194f9bull, 0xab1c5ed5da6d8118ull,
    0xd807aa98a3030242ull, 0x12835b0145706fbeull, 0x243185be4ee4b28cull, 0x550c7dc3d5ffb4e2ull, 0x72be5d74f27b896full, 0x80deb1fe3b1696b1ull, 0x9bdc06a725c71235ull, 0xc19bf174cf692694ull,
    0xe49b69c19ef14ad2ull, 0xefbe4786384f25e3ull, 0x0fc19dc68b8cd5b5ull, 0x240ca1cc77ac9c65ull, 0x2de92c6f592b0275ull, 0x4a7484aa6ea6e483ull, 0x5cb0a9dcbd41fbd4ull, 0x76f988da831153b5ull,
    0x983e5152ee66dfabull, 0xa831c66d2db43210ull, 0xb00327c898fb213full, 0xbf597fc7beef0ee4ull, 0xc6e00bf33da88fc2ull, 0xd5a79147930aa725ull, 0x06ca6351e003826full, 0x142929670a0e6e70ull,
    0x27b70a8546d22ffcull, 0x2e1b21385c26c926ull, 0x4d2c6dfc5ac42aedull, 0x53380d139d95b3dfull, 0x650a73548baf63deull, 0x766a0abb3c77b2a8ull, 0x81c2c92e47edaee6ull, 0x92722c851482353bull,
    0xa2bfe8a14cf10364ull, 0xa81a664bbc423001ull, 0xc24b8b70d0f89791ull, 0xc76c51a30654be30ull, 0xd192e819d6ef5218ull, 0xd69906245565a910ull, 0xf40e35855771202aull, 0x106aa07032bbd1b8ull,
    0x19a4c116b8d2d0c8ull, 0x1e376c085141ab53ull, 0x2748774cdf8eeb99ull, 0x34b0bcb5e19b48a8ull, 0x391c0cb3c5c95a63ull, 0x4ed8aa4ae3418acbull, 0x5b9cca4f7763e373ull, 0x682e6ff3d6b2b8a3ull,
    0x748f82ee5defb2fcull, 0x78a5636f43172f60ull, 0x84c87814a1f0ab72ull, 0x8cc702081a6439ecull, 0x90befffa23631e28ull, 0xa4506cebde82bde9ull, 0xbef9a3f7b2c67915ull, 0xc67178f2e372532bull,
    0xca273eceea26619cull, 0xd186b8c721c0c207ull, 0xeada7dd6cde0eb1eull, 0xf57d4f7fee6ed178ull, 0x06f067aa72176fbaull, 0x0a637dc5a2c898a6ull, 0x113f9804bef90daeull, 0x1b710b35131c471bull,
    0x28db77f523047d84ull, 0x32caab7b40c72493ull, 0x3c9ebe0a15c9bebcull, 0x431d67c49c100d4cull, 0x4cc5d4becb3e42b6ull, 0x597f299cfc657e2aull, 0x5fcb6fab3ad6faecull, 0x6c44198c4a475817ull};
__device__ __forceinline__ uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    return ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
}
// first 32 bytes of SHA-512(a LE || b LE) as four little-endian 64-bit limbs
__device__ void sha512_two_words(uint64_t a, uint64_t b, uint64_t out[4]) {
    uint64_t w[16];
    w[0] = bswap64(a); w[1] = bswap64(b); w[2] = 0x8000000000000000ull;
    for (int i = 3; i < 15; i++) w[i] = 0;
    w[15] = 128;                                             // message length in bits
    uint64_t h[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                     0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    uint64_t s[8];
    for (int i = 0; i < 8; i++) s[i] = h[i];
    for (int t = 0; t < 80; t++) {
        if (t >= 16) {
            const uint64_t w15 = w[(t - 15) & 15], w2 = w[(t - 2) & 15];
            const uint64_t s0 = rotr64(w15, 1) ^ rotr64(w15, 8) ^ (w15 >> 7), s1 = rotr64(w2, 19) ^ rotr64(w2, 61) ^ (w2 >> 6);
            w[t & 15] = w[t & 15] + s0 + w[(t - 7) & 15] + s1;
        }
        const uint64_t S1 = rotr64(s[4], 14) ^ rotr64(s[4], 18) ^ rotr64(s[4], 41), ch = (s[4] & s[5]) ^ (~s[4] & s[6]);
        const uint64_t t1 = s[7] + S1 + ch + SHA512_K[t] + w[t & 15];
        const uint64_t S0 = rotr64(s[0], 28) ^ rotr64(s[0], 34) ^ rotr64(s[0], 39), maj = (s[0] & s[1]) ^ (s[0] & s[2]) ^ (s[1] & s[2]);
        const uint64_t t2 = S0 + maj;
        s[7] = s[6]; s[6] = s[5]; s[5] = s[4]; s[4] = s[3] + t1; s[3] = s[2]; s[2] = s[1]; s[1] = s[0]; s[0] = t1 + t2;
    }
    for (int i = 0; i < 4; i++) out[i] = bswap64(h[i] + s[i]);   // digest bytes are big-endian words; the bigint reads the bytes little-endian
}
__device__ void sha512_rng_fr(uint64_t idx, uint32_t out[8]) {
    for (uint64_t iter = 0;; iter++) {
        uint64_t v[4];
        sha512_two_words(idx, iter, v);
        v[3] &= 0x3fffffffffffffffull;                       // bits >= 254 cleared (the modulus has 254 bits)
        Fr x; for (int i = 0; i < 4; i++) { x.v[2 * i] = (uint32_t)v[i]; x.v[2 * i + 1] = (uint32_t)(v[i] >> 32); }
        Fr y = x; y.reduce_once();
        if (y == x) { for (int i = 0; i < 8; i++) out[i] = x.v[i]; return; }      // below r: accepted
    }
}
__global__ void fill_scalars_kernel(uint32_t *out, size_t n, size_t first = 0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sha512_rng_fr(first + i, out + i * 8);
}
// table[k*15 + d-1] = d * 16^k * G (affine), k < 64, d = 1..15: one thread per k
template <class F> __global__ void gen_table_kernel(Affine<F> gen, Affine<F> *table) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 64) return;
    XYZZ<F> base = XYZZ<F>::from_affine(gen);
    for (int j = 0; j < 4 * k; j++) base = base.dbl();
    const Affine<F> b = base.to_affine();
    XYZZ<F> cur = XYZZ<F>::from_affine(b);
    for (int d = 1; d <= 15; d++) { table[k * 15 + d - 1] = cur.to_affine(); cur.add_affine(b); }
}
template <class F> __global__ void gen_bases_kernel(const Affine<F> *__restrict__ table, Affine<F> *out, size_t n, size_t first = 0) {   // P_i = SHA512_rng(2^32 + i) * G
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h[8];
    sha512_rng_fr((1ull << 32) + first + i, h);
    XYZZ<F> r = XYZZ<F>::inf();
    for (int k = 0; k < 64; k++) {
        const uint32_t d = (h[k >> 3] >> ((k & 7) * 4)) & 15;
        if (d) r.add_affine(table[k * 15 + d - 1]);
    }
    out[i] = r.to_affine();
}

// Dependent-free mad.lo.u32 streams (8 independent accumulators per thread).  SASS check: the `wide == 0` path is a run of
// `IMAD Rk, Ra, Rb, Rk`; ptxas strength-reduces the loop-invariant mad.wide variant to IADD3, so only wide == 0 is a valid peak.
// Integer-multiply peaks of the device (roofline denominators; same kernels as scripts/ubench/imad.cu):
//   mode 0  32x32->32 multiply-add (IMAD),        dependent-free mad.lo.u32 chains
//   mode 1  32x32->64 multiply-add (IMAD.WIDE.X), carry-chained mad.lo.cc/madc.hi.cc pairs exactly as the field multiplication
//           issues them (ptxas fuses each pair into one IMAD.WIDE.U32.X)
//   mode 2  whole 254-bit Montgomery multiplications (Fq), one dependent chain per thread
__global__ void imad_peak_kernel(uint32_t *out, int iters, int mode) {
    const uint32_t a = threadIdx.x * 2654435761u + 1, b = blockIdx.x * 40503u + 3;
    uint32_t r = 0;
    if (mode == 0) {
        uint32_t x[8] = {a, a * 3, a ^ b, a + b, a * 5, a * 7, a * 11, a * 13};
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(x[k]) : "r"(b));
        }
        for (int k = 0; k < 8; k++) r ^= x[k];
    } else if (mode == 1) {
        uint32_t v[18];
        for (int k = 0; k < 18; k++) v[k] = a * (k + 1);
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(v[0]) : "r"(v[16]), "r"(b));
                asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(v[1]) : "r"(v[16]), "r"(b));
#pragma unroll
                for (int k = 2; k < 16; k += 2) {
                    asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(v[k]) : "r"(v[17]), "r"(b));
                    asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(v[k + 1]) : "r"(v[17]), "r"(b));
                }
                asm volatile("addc.u32 %0, %0, 0;" : "+r"(v[16]));
                v[17] ^= v[3];
            }
        }
        for (int k = 0; k < 18; k++) r ^= v[k];
    } else {
        Fq x = Fq::one(), y;
        for (int k = 0; k < 8; k++) y.v[k] = Fq::r2().v[k] ^ (threadIdx.x & 0xff);
        y.v[7] &= 0x0fffffff;
        x.v[0] += blockIdx.x;
        for (int i = 0; i < iters * 64; i++) x = Fq::mul_impl(x, y);
        for (int k = 0; k < 8; k++) r ^= x.v[k];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// every exported function below is declared extern "C" in include/zkb200.h, which fixes its linkage

int zkb200_init(int device) {
    const int rc = init_device(device);
    if (rc == 0) g_explicit = true;
    return rc;
}
int zkb200_ensure_device(void) { return ensure_device() ? -1 : g_device; }
int zkb200_device_count(void) { int count = 0; return cudaGetDeviceCount(&count) == cudaSuccess ? count : 0; }
int zkb200_current_device(void) { return g_explicit ? g_device : -1; }
const char *zkb200_last_error(void) { std::lock_guard<std::mutex> lk(g_err_mu); static thread_local std::string copy; copy = g_err; return copy.c_str(); }

void *zkb200_pk_load_on(const char *path, int device) {
    if (device < 0 || device >= zkb200_device_count()) { set_err("bad device index"); return nullptr; }
    std::string err;
    DevicePk *pk = pk_load(path, device, err);
    if (!pk) set_err(err);
    return pk;
}
void *zkb200_pk_load(const char *path) {
    if (ensure_device()) return nullptr;
    return zkb200_pk_load_on(path, g_device);
}
// one parse of the key file, one resident copy per device, uploaded by one host thread per device
int zkb200_pk_load_many(const char *path, const int *devices, int n, void **out) {
    if (n <= 0 || !devices || !out) return -1;
    const int count = zkb200_device_count();
    for (int i = 0; i < n; i++) { out[i] = nullptr; if (devices[i] < 0 || devices[i] >= count) { set_err("bad device index"); return -1; } }
    std::string err; double parse_s = 0;
    zkpk::ParsedPk *P = pk_parse_file(path, err, &parse_s);
    if (!P) { set_err(err); return -1; }
    std::vector<std::string> errs(n);
    std::vector<std::thread> th;
    for (int i = 1; i < n; i++) th.emplace_back([&, i]() { out[i] = pk_from_parsed(*P, devices[i], errs[i], parse_s); });
    out[0] = pk_from_parsed(*P, devices[0], errs[0], parse_s);
    for (auto &t : th) t.join();
    pk_parsed_free(P);
    int rc = 0;
    for (int i = 0; i < n; i++) if (!out[i]) { set_err(errs[i]); rc = -1; }
    if (rc) for (int i = 0; i < n; i++) if (out[i]) { pk_free((DevicePk *)out[i]); out[i] = nullptr; }
    return rc;
}
void zkb200_pk_free(void *pk) { pk_free((DevicePk *)pk); }
int zkb200_pk_info(void *h, uint64_t info[8], double seconds[3]) {
    DevicePk *pk = (DevicePk *)h;
    if (!pk) return -1;
    info[0] = pk->num_vars; info[1] = pk->num_inputs; info[2] = pk->num_constraints; info[3] = pk->dom->m; info[4] = pk->dom->step ? 1 : 0;
    info[5] = (uint64_t)pk->a.nnz + pk->b.nnz + pk->c.nnz; info[6] = pk->ncoef; info[7] = pk->nB;
    if (seconds) { seconds[0] = pk->load_seconds; seconds[1] = pk->parse_seconds; seconds[2] = pk->decompress_seconds + pk->expand_seconds; }
    return 0;
}

static void put_fq(uint8_t *o, const zkh::HFq &x) { uint64_t c[4]; x.to_canonical(c); memcpy(o, c, 32); }
static void put_g1(uint8_t *o, const zkh::HG1Affine &a) { if (a.is_inf()) { memset(o, 0, 64); return; } put_fq(o, a.x); put_fq(o + 32, a.y); }
static void put_g2(uint8_t *o, const zkh::HG2Affine &a) {
    if (a.is_inf()) { memset(o, 0, 128); return; }
    put_fq(o, a.x.c0); put_fq(o + 32, a.x.c1); put_fq(o + 64, a.y.c0); put_fq(o + 96, a.y.c1);
}

int zkb200_pk_device(void *h) { return h ? ((DevicePk *)h)->device : -1; }

static const char *DEFAULT_PROOF =   // (G1::one, G2::one, G1::one): r1cs_gg_ppzksnark_proof default ctor (r1cs_gg_ppzksnark.hpp:309-315)
    "0000000000000000000000000000000000000000000000000000000000000001"
    "0000000000000000000000000000000000000000000000000000000000000002"
    "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2"
    "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"
    "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"
    "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa"
    "0000000000000000000000000000000000000000000000000000000000000001"
    "0000000000000000000000000000000000000000000000000000000000000002";

const char *zkb200_default_proof(void) { return DEFAULT_PROOF; }

// proof bytes, parity hooks and timings of a collected proof
static int finish_outputs(ProofPoints &pp, double host_ms, char *proof_hex_out, uint8_t *parts, float *timings_ms) {
    if (parts) { put_g1(parts, pp.At); put_g2(parts + 64, pp.Bt_g); put_g1(parts + 192, pp.Bt_h); put_g1(parts + 256, pp.Ht); put_g1(parts + 320, pp.Lt); }
    if (timings_ms) { timings_ms[0] = pp.gpu_ms; timings_ms[1] = pp.qap_ms; timings_ms[2] = pp.msm_h_ms; timings_ms[3] = (float)host_ms; timings_ms[4] = pp.acc_h_ms;
                      timings_ms[5] = pp.a_done_ms; timings_ms[6] = pp.b_done_ms; timings_ms[7] = pp.l_done_ms; }
    const std::string hex = pp.satisfied ? proof_hex(pp) : std::string(DEFAULT_PROOF);
    memcpy(proof_hex_out, hex.data(), 512); proof_hex_out[512] = 0;
    return pp.satisfied ? 0 : 1;
}
static int prove_any(void *h, const uint8_t *assignment, const uint64_t *lo, const WideIn *wide, uint32_t nwide, const uint8_t r[32], const uint8_t s[32],
                     char *proof_hex_out, uint8_t *parts, float *timings_ms, const void *seeds = nullptr, uint32_t nseeds = 0) {
    DevicePk *pk = (DevicePk *)h;
    if (!pk) return -1;
    uint64_t rr[4], ss[4]; memcpy(rr, r, 32); memcpy(ss, s, 32);
    ProofPoints pp;
    pp.want_parts = parts != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = lo ? prove_compact(pk, lo, wide, nwide, rr, ss, pp, seeds, nseeds) : prove(pk, assignment, rr, ss, pp);
    (void)t0;
    if (rc < 0) { set_err("prove: malformed input (" + std::to_string(rc) + ")"); return rc; }
    return finish_outputs(pp, pp.host_tail_ms, proof_hex_out, parts, timings_ms);
}
int zkb200_prove(void *h, const uint8_t *assignment, const uint8_t r[32], const uint8_t s[32], char *proof_hex_out, uint8_t *parts, float *timings_ms) {
    return prove_any(h, assignment, nullptr, nullptr, 0, r, s, proof_hex_out, parts, timings_ms);
}
int zkb200_prove_compact(void *h, const uint64_t *lo, const void *wide, size_t nwide, const void *seeds, size_t nseeds, const uint8_t r[32], const uint8_t s[32],
                         char *proof_hex_out, float *timings_ms) {
    if (nseeds > 0xffffffffull) return -5;
    return prove_any(h, nullptr, lo, (const WideIn *)wide, (uint32_t)nwide, r, s, proof_hex_out, nullptr, timings_ms, seeds, (uint32_t)nseeds);
}
int zkb200_lane_read_assignment(void *h, int lane, uint8_t *out) {
    DevicePk *pk = (DevicePk *)h;
    if (!pk || lane < 0 || lane >= pk->nlanes) return -1;
    return lane_read_assignment(pk, pk->lanes[lane], out);
}
int zkb200_pk_lanes(void *h) { return h ? ((DevicePk *)h)->nlanes : 0; }
int zkb200_lane_acquire(void *h) { return h ? lane_acquire((DevicePk *)h)->index : -1; }
static Lane *held_lane(void *h, int lane) {
    DevicePk *pk = (DevicePk *)h;
    if (!pk || lane < 0 || lane >= pk->nlanes || !pk->lanes[lane]->busy) return nullptr;
    return pk->lanes[lane];
}
void zkb200_lane_release(void *h, int lane) { if (Lane *ln = held_lane(h, lane)) lane_release((DevicePk *)h, ln); }
uint64_t *zkb200_lane_staging(void *h, int lane) { Lane *ln = held_lane(h, lane); return ln ? compact_staging(ln) : nullptr; }
int zkb200_prove_submit(void *h, int lane, const uint8_t *assignment, const uint8_t r[32], const uint8_t s[32]) {
    Lane *ln = held_lane(h, lane);
    if (!ln || ln->pending) return -1;
    uint64_t rr[4], ss[4]; memcpy(rr, r, 32); memcpy(ss, s, 32);
    return prove_submit((DevicePk *)h, ln, assignment, nullptr, nullptr, 0, rr, ss);
}
int zkb200_prove_submit_compact(void *h, int lane, const uint64_t *lo, const void *wide, size_t nwide, const void *seeds, size_t nseeds, const uint8_t r[32],
                                const uint8_t s[32]) {
    Lane *ln = held_lane(h, lane);
    if (!ln || ln->pending || !lo) return -1;
    if (nseeds > 0xffffffffull) return -5;
    uint64_t rr[4], ss[4]; memcpy(rr, r, 32); memcpy(ss, s, 32);
    return prove_submit((DevicePk *)h, ln, nullptr, lo, (const WideIn *)wide, (uint32_t)nwide, rr, ss, seeds, (uint32_t)nseeds);
}
int zkb200_prove_collect(void *h, int lane, char *proof_hex_out, uint8_t *parts, float *timings_ms) {
    Lane *ln = held_lane(h, lane);
    if (!ln || !ln->pending) return -1;
    ProofPoints pp;
    pp.want_parts = parts != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (prove_collect((DevicePk *)h, ln, pp)) return -1;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    (void)ms;
    return finish_outputs(pp, pp.host_tail_ms, proof_hex_out, parts, timings_ms);
}
int zkb200_qap_witness_map(void *h, const uint8_t *assignment, uint8_t *out_H, int *satisfied) {
    DevicePk *pk = (DevicePk *)h;
    if (!pk) return -1;
    return qap_witness_map(pk, assignment, out_H, satisfied);
}
int zkb200_last_launches(void) { return launches_last_prove(); }
void zkb200_last_transfer_bytes(unsigned long long out[2]) { transfer_bytes_last_prove(out); }
void zkb200_set_isolate_h(int on) { set_isolate_h(on != 0); }

// ---- evaluation domains ---------------------------------------------------------------------------------------------------
static std::map<std::pair<int, uint64_t>, Domain *> g_domains;
static Domain *get_domain(uint64_t min_size) {
    auto key = std::make_pair(g_device, min_size);
    auto it = g_domains.find(key);
    if (it != g_domains.end()) return it->second;
    Domain *d = Domain::build(min_size);
    g_domains[key] = d;
    return d;
}
long zkb200_domain_op(size_t min_size, int op, uint8_t *data, size_t n, int *kind) {
    if (ensure_device()) return -1;
    std::lock_guard<std::mutex> lk(g_mu);
    Domain *d = get_domain(min_size);
    if (!d) { set_err("no evaluation domain for that size"); return -1; }
    if (kind) *kind = d->step ? 1 : 0;
    if (!data) return d->m;
    if (n != d->m || op < 0 || op > 4) { set_err("domain_op: wrong vector length or op"); return -1; }
    Fr *buf, *tmp;
    ZK_CUDA(cudaMalloc(&buf, n * 32)); ZK_CUDA(cudaMalloc(&tmp, n * 32));
    ZK_CUDA(cudaMemcpy(buf, data, n * 32, cudaMemcpyHostToDevice));
    to_mont_kernel<<<(unsigned)((n + 255) / 256), 256>>>(buf, (uint32_t)n);
    domain_op(0, *d, op, buf, tmp);
    from_mont_kernel<<<(unsigned)((n + 255) / 256), 256>>>(buf, tmp, (uint32_t)n);
    ZK_CUDA(cudaMemcpy(data, tmp, n * 32, cudaMemcpyDeviceToHost));
    ZK_CUDA(cudaFree(buf)); ZK_CUDA(cudaFree(tmp));
    return d->m;
}

// ---- MSM -------------------------------------------------------------------------------------------------------------------
static int default_window(size_t n) {
    int c = 0; while ((1ull << c) < n) c++;
    c -= 4; if (c < 4) c = 4; if (c > 16) c = 16;
    return c;
}
template <class AffT, class HP>
static int msm_host_entry(size_t n, const uint8_t *bases, const uint8_t *scalars, int window_bits, bool g2, HP &result) {
    if (ensure_device()) return -1;
    std::lock_guard<std::mutex> lk(g_mu);
    const int c = window_bits > 0 ? window_bits : default_window(n);
    AffT *d_bases; uint32_t *d_scalars;
    ZK_CUDA(cudaMalloc(&d_bases, (n + 1) * sizeof(AffT))); ZK_CUDA(cudaMalloc(&d_scalars, (n + 1) * 32));
    ZK_CUDA(cudaMemcpy(d_bases, bases, n * sizeof(AffT), cudaMemcpyHostToDevice));
    ZK_CUDA(cudaMemcpy(d_scalars, scalars, n * 32, cudaMemcpyHostToDevice));
    const size_t nf = n * sizeof(AffT) / 32;
    if (nf) to_mont_generic_kernel<Fq><<<(unsigned)((nf + 255) / 256), 256>>>((Fq *)d_bases, nf);
    MsmPlan plan; plan.init((uint32_t)n, c, 1024, !g2, g2, false);
    msm_run(0, plan, ScalarRef{d_scalars, nullptr, 0, 0}, nullptr, g2 ? nullptr : d_bases, g2 ? d_bases : nullptr);
    ZK_CUDA(cudaStreamSynchronize(0));
    if constexpr (sizeof(AffT) == 64) result = msm_finish_g1(plan); else result = msm_finish_g2(plan);
    plan.release();
    ZK_CUDA(cudaFree(d_bases)); ZK_CUDA(cudaFree(d_scalars));
    return 0;
}
int zkb200_msm_g1(size_t n, const uint8_t *bases, const uint8_t *scalars, int window_bits, uint8_t out[64]) {
    zkh::HG1 r;
    if (msm_host_entry<G1Affine>(n, bases, scalars, window_bits, false, r)) return -1;
    put_g1(out, r.to_affine());
    return 0;
}
int zkb200_msm_g2(size_t n, const uint8_t *bases, const uint8_t *scalars, int window_bits, uint8_t out[128]) {
    zkh::HG2 r;
    if (msm_host_entry<G2Affine>(n, bases, scalars, window_bits, true, r)) return -1;
    put_g2(out, r.to_affine());
    return 0;
}

int zkb200_field_op(int field, int op, size_t n, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    if (ensure_device()) return -1;
    std::lock_guard<std::mutex> lk(g_mu);
    void *da, *db = nullptr, *dout;
    ZK_CUDA(cudaMalloc(&da, n * 32)); ZK_CUDA(cudaMalloc(&dout, n * 32));
    ZK_CUDA(cudaMemcpy(da, a, n * 32, cudaMemcpyHostToDevice));
    if (b) { ZK_CUDA(cudaMalloc(&db, n * 32)); ZK_CUDA(cudaMemcpy(db, b, n * 32, cudaMemcpyHostToDevice)); }
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (field == 0) field_op_kernel<Fr><<<grid, 128>>>((const Fr *)da, (const Fr *)db, (Fr *)dout, n, op);
    else field_op_kernel<Fq><<<grid, 128>>>((const Fq *)da, (const Fq *)db, (Fq *)dout, n, op);
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpy(out, dout, n * 32, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(dout); if (db) cudaFree(db);
    return 0;
}

// ---- device-resident benchmarks ------------------------------------------------------------------------------------------
float zkb200_bench_ntt(int logn, int batch, int iters) {
    if (ensure_device()) return -1;
    if (logn < 1 || logn > 28 || batch < 1) return -1;
    Domain *d = get_domain(1ull << logn);
    const size_t n = 1ull << logn;
    Fr *src, *dst;
    ZK_CUDA(cudaMalloc(&src, n * 32 * batch)); ZK_CUDA(cudaMalloc(&dst, n * 32 * batch));
    fill_scalars_kernel<<<(unsigned)((n * batch + 255) / 256), 256>>>((uint32_t *)src, n * batch, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const PowMul none{nullptr, nullptr, 0};
    for (int b = 0; b < batch; b++) ntt_launch(0, src + b * n, dst + b * n, (const Fr *)d->tw_big_f, logn, none, none);   // warm-up
    zkp::device_sync();
    cudaEventRecord(e0, 0);
    for (int it = 0; it < iters; it++)
        for (int b = 0; b < batch; b++) ntt_launch(0, src + b * n, dst + b * n, (const Fr *)d->tw_big_f, logn, none, none);
    cudaEventRecord(e1, 0);
    ZK_CUDA(cudaEventSynchronize(e1));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(src); cudaFree(dst);
    return ms / (float)(iters * batch);
}

template <class F> __global__ void from_mont_generic_kernel(F *a, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = a[i].from_mont();
}
// device array of n (+1 spare) synthetic bases P_{first+i}, Montgomery form; group 1 = G1, 2 = G2
static void *synth_bases_device(int group, size_t first, size_t n) {
    void *bases, *table;
    if (group == 1) {
        ZK_CUDA(cudaMalloc(&bases, (n + 1) * sizeof(G1Affine))); ZK_CUDA(cudaMalloc(&table, 64 * 15 * sizeof(G1Affine)));
        G1Affine g; g.x = Fq::one(); g.y = Fq::one() + Fq::one();
        gen_table_kernel<Fq><<<1, 64>>>(g, (G1Affine *)table);
        if (n) gen_bases_kernel<Fq><<<(unsigned)((n + 127) / 128), 128>>>((const G1Affine *)table, (G1Affine *)bases, n, first);
    } else {
        ZK_CUDA(cudaMalloc(&bases, (n + 1) * sizeof(G2Affine))); ZK_CUDA(cudaMalloc(&table, 64 * 15 * sizeof(G2Affine)));
        // G2 generator (alt_bn128_init.cpp:265-269), Montgomery form computed on the host
        const char *gs[4] = {"10857046999023057135944570762232829481370756359578518086990519993285655852781",
                             "11559732032986387107991004021392285783925812861821192530917403151452391805634",
                             "8495653923123431417604973247489272438418190587263600148770280649306958101930",
                             "4082367875863433681332203403145435568316851327593401208105741076214120093531"};
        zkh::HFq v[4]; for (int i = 0; i < 4; i++) zkh::HFq::from_dec(gs[i], strlen(gs[i]), v[i]);
        G2Affine g; memcpy(&g, v, 128);
        gen_table_kernel<Fq2><<<1, 64>>>(g, (G2Affine *)table);
        if (n) gen_bases_kernel<Fq2><<<(unsigned)((n + 127) / 128), 128>>>((const G2Affine *)table, (G2Affine *)bases, n, first);
    }
    ZK_CUDA(cudaGetLastError());
    zkp::device_sync();
    cudaFree(table);
    return bases;
}
// the synthetic inputs themselves, for the host legs of the sweep and the parity tests: canonical little-endian bytes
int zkb200_synth_scalars(size_t first, size_t n, uint8_t *out) {
    if (ensure_device()) return -1;
    uint32_t *sc; ZK_CUDA(cudaMalloc(&sc, (n + 1) * 32));
    if (n) fill_scalars_kernel<<<(unsigned)((n + 255) / 256), 256>>>(sc, n, first);
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpy(out, sc, n * 32, cudaMemcpyDeviceToHost));
    cudaFree(sc);
    return 0;
}
int zkb200_synth_bases(int group, size_t first, size_t n, uint8_t *out) {
    if (ensure_device() || (group != 1 && group != 2)) return -1;
    void *bases = synth_bases_device(group, first, n);
    const size_t nf = n * (group == 1 ? 2 : 4);
    if (nf) from_mont_generic_kernel<Fq><<<(unsigned)((nf + 255) / 256), 256>>>((Fq *)bases, nf);
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpy(out, bases, nf * 32, cudaMemcpyDeviceToHost));
    cudaFree(bases);
    return 0;
}

// slice [first, first+n) of the synthetic MSM problem (bases P_i, scalars s_i are functions of the GLOBAL index i), so that a single
// large MSM can be split by point range over several GPUs: every rank calls this with its slice and the partial points are added.
float zkb200_bench_msm_slice(int group, size_t first, size_t n, int window_bits, int iters, uint8_t *out_point) {
    if (ensure_device()) return -1;
    const int c = window_bits > 0 ? window_bits : (window_bits < 0 ? -window_bits : default_window(n));
    uint32_t *sc; ZK_CUDA(cudaMalloc(&sc, (n + 1) * 32));
    if (n) fill_scalars_kernel<<<(unsigned)((n + 255) / 256), 256>>>(sc, n, first);
    void *bases = synth_bases_device(group, first, n);
    zkp::device_sync();
    const bool expanded = window_bits < 0;          // negative window_bits: fixed-base (expanded) layout with |window_bits| bits
    int aff = 0;                                     // affine halving rounds in front of the accumulation: off unless asked for, as in the prover
    if (const char *e = getenv("ZKB200_AFFINE_ROUNDS")) aff = atoi(e);
    MsmPlan plan; plan.init((uint32_t)n, c, 0, group == 1, group == 2, expanded, expanded ? aff : 0);
    if (expanded) { void *e = msm_expand_bases(bases, (uint32_t)n, c, group == 2); zkp::device_sync(); cudaFree(bases); bases = e; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    msm_run(0, plan, ScalarRef{sc, nullptr, 0, 0}, nullptr, group == 1 ? bases : nullptr, group == 2 ? bases : nullptr);
    zkp::device_sync();
    cudaEventRecord(e0, 0);
    for (int it = 0; it < iters; it++)
        msm_run(0, plan, ScalarRef{sc, nullptr, 0, 0}, nullptr, group == 1 ? bases : nullptr, group == 2 ? bases : nullptr);
    cudaEventRecord(e1, 0);
    ZK_CUDA(cudaEventSynchronize(e1));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (out_point) { if (group == 1) put_g1(out_point, msm_finish_g1(plan).to_affine()); else put_g2(out_point, msm_finish_g2(plan).to_affine()); }
    plan.release(); cudaFree(sc); cudaFree(bases);
    return ms / (float)(iters > 0 ? iters : 1);
}
// host: sum of n affine G1 points (64 B each, all-zero = infinity) -- the last step of an MSM split by point range over several GPUs
int zkb200_g1_sum(size_t n, const uint8_t *points, uint8_t out[64]) {
    zkh::HG1 acc = zkh::HG1::inf();
    for (size_t i = 0; i < n; i++) {
        uint64_t c[8]; memcpy(c, points + 64 * i, 64);
        if (!(c[0] | c[1] | c[2] | c[3] | c[4] | c[5] | c[6] | c[7])) continue;
        zkh::HG1Affine a{zkh::HFq::from_canonical(c), zkh::HFq::from_canonical(c + 4)};
        acc = acc.add(zkh::HG1::from_affine(a));
    }
    put_g1(out, acc.to_affine());
    return 0;
}
int zkb200_g2_sum(size_t n, const uint8_t *points, uint8_t out[128]) {
    zkh::HG2 acc = zkh::HG2::inf();
    for (size_t i = 0; i < n; i++) {
        uint64_t c[16]; memcpy(c, points + 128 * i, 128);
        uint64_t any = 0; for (int k = 0; k < 16; k++) any |= c[k];
        if (!any) continue;
        zkh::HG2Affine a{zkh::HFq2{zkh::HFq::from_canonical(c), zkh::HFq::from_canonical(c + 4)}, zkh::HFq2{zkh::HFq::from_canonical(c + 8), zkh::HFq::from_canonical(c + 12)}};
        acc = acc.add(zkh::HG2::from_affine(a));
    }
    put_g2(out, acc.to_affine());
    return 0;
}
float zkb200_bench_msm(int group, size_t n, int window_bits, int iters) { return zkb200_bench_msm_slice(group, 0, n, window_bits, iters, nullptr); }

// write a buffer twice the size of L2 on every device in use so the next step starts with a cold L2 (bench hygiene)
void zkb200_flush_l2(void) {
    if (ensure_device()) return;
    static void *buf[64];
    static int flip = 0;
    const size_t bytes = 256u << 20;
    ++flip;
    for (int d : devices_in_use()) {
        ZK_CUDA(cudaSetDevice(d));
        if (!buf[d]) ZK_CUDA(cudaMalloc(&buf[d], bytes));
        ZK_CUDA(cudaMemsetAsync(buf[d], flip, bytes, 0));
    }
    zkb200_device_sync();
}
void zkb200_device_sync(void) {
    if (ensure_device()) return;
    for (int d : devices_in_use()) { ZK_CUDA(cudaSetDevice(d)); zkp::device_sync(); }
    cudaSetDevice(g_device);
}
// Device-clock stopwatch around a region that runs on many streams (and, in a single process driving several GPUs, on many devices): both
// calls synchronise every device in use first and then record a CUDA event on each, so the elapsed time is GPU time between two quiescent
// points.  stop = 0 starts; stop = 1 returns the milliseconds since the start (the longest over the devices).
float zkb200_device_timer(int stop) {
    if (ensure_device()) return -1;
    static cudaEvent_t ev[64][2];
    float worst = 0;
    zkb200_device_sync();
    for (int d : devices_in_use()) {
        ZK_CUDA(cudaSetDevice(d));
        if (!ev[d][0]) { ZK_CUDA(cudaEventCreate(&ev[d][0])); ZK_CUDA(cudaEventCreate(&ev[d][1])); }
        ZK_CUDA(cudaEventRecord(ev[d][stop ? 1 : 0], 0));
        ZK_CUDA(cudaEventSynchronize(ev[d][stop ? 1 : 0]));
        if (stop) { float ms = 0; if (cudaEventElapsedTime(&ms, ev[d][0], ev[d][1]) == cudaSuccess && ms > worst) worst = ms; else cudaGetLastError(); }
    }
    cudaSetDevice(g_device);
    return worst;
}

float zkb200_bench_imad_peak(int mode) {
    if (ensure_device()) return -1;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, g_device);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = mode == 2 ? 32 : 4096;
    uint32_t *out; ZK_CUDA(cudaMalloc(&out, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    imad_peak_kernel<<<blocks, threads>>>(out, 2, mode);
    zkp::device_sync();
    cudaEventRecord(e0, 0);
    imad_peak_kernel<<<blocks, threads>>>(out, iters, mode);
    cudaEventRecord(e1, 0);
    ZK_CUDA(cudaEventSynchronize(e1));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    // per thread and iteration: 64 IMAD (mode 0), 64 wide multiply-adds (mode 1), 64 modular multiplications (mode 2)
    const double ops = (double)blocks * threads * iters * 64.0;
    return (float)(ops / (ms * 1e-3) / 1e12);             // tera-ops per second
}


