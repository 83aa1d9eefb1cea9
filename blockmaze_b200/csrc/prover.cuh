// Device-side state of the B200 Groth16 prover: evaluation domains, resident proving keys, MSM work areas.
// Declarations shared by prover.cu (GPU pipeline) and cabi.cpp (C-ABI).  See DESIGN.md for the data layout in HBM.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
#include "host_field.hpp"

namespace zkpk { struct ParsedPk; }

namespace zkp {

void cuda_check(cudaError_t e, const char *what);     // aborts loudly: there is no CPU fallback
#define ZK_CUDA(x) ::zkp::cuda_check((x), #x)

// ---- evaluation domain (libfqfft get_evaluation_domain: basic_radix2 or step_radix2) ------------------------------
struct Domain {
    uint32_t m = 0, big = 0, small = 0;      // basic: big = m, small = 0
    int log_big = 0, log_small = 0;
    bool step = false;
    uint32_t compr = 1;
    // twiddles (device, Montgomery): per-stage tables (ntt.cuh ntt_tw) of the two power-of-two sub-transforms, forward and inverse
    void *tw_big_f = nullptr, *tw_big_i = nullptr, *tw_small_f = nullptr, *tw_small_i = nullptr;
    // step only: omega^j / omega^-j, j < big, omega = primitive (2*big)-th root
    void *tw_step_f = nullptr, *tw_step_i = nullptr;
    // two-level power tables of the coset generator g = 5: lo[i & 1023], hi[i >> 10]
    void *g_lo = nullptr, *g_hi = nullptr;              // g^i
    void *g_hi_ninv = nullptr;                          // g^i / m        (basic: iFFT scale folded into the coset shift)
    void *gi_lo = nullptr, *gi_hi = nullptr;            // g^-i
    void *gi_hi_ninv = nullptr;                         // g^-i / m
    void *gi_hi_zninv = nullptr, *c_z_over_m = nullptr; // basic: g^-i / (m Z(g)) and the constant 1 / (m Z(g)) (six-transform QAP map)
    void *g_full_ninv = nullptr, *gi_full_zninv = nullptr;   // basic: g^i / m and g^-i / (m Z(g)) for every i < m (one multiplication per element instead of two)
    void *c_big_inv = nullptr, *c_small_inv = nullptr, *c_m_inv = nullptr;   // single constants 1/big, 1/small, 1/m
    void *zt = nullptr;                                 // 1/Z on the coset, `compr` distinct values for i < big
    zkh::HFr z1, over_two;                              // 1/Z for i >= big (step); 1/2
    static Domain *build(uint64_t min_size);            // selection rule of get_evaluation_domain.tcc:33-52
    void release();
};

enum DomainOp { OP_FFT = 0, OP_IFFT = 1, OP_COSET_FFT = 2, OP_ICOSET_FFT = 3, OP_DIVIDE_BY_Z = 4 };
// in place on device buffer `data` (m Montgomery elements); `tmp` is scratch of the same size
void domain_op(cudaStream_t st, const Domain &d, int op, void *data, void *tmp);

// ---- MSM work area ---------------------------------------------------------------------------------------------------
struct MsmPlan {
    uint32_t n = 0;             // number of bases
    int c = 0, windows = 0;
    uint32_t nb = 0, ones = 0, total = 0;
    uint32_t seg = 0, seg_weighted = 0, bpw = 0;  // reduce: buckets per thread (ones / weighted regions), CTAs per window
    void *counts = nullptr, *offsets = nullptr, *cursors = nullptr, *entries = nullptr;
    size_t entries_cap = 0;
    bool sort_v2_now = false;   // the run in flight used the absolute-cursor sort (its scan cleared counts and heavy queues)
    bool expanded = false;      // bases hold 2^(c*k)*P for every window k: one bucket region, no Horner
    uint32_t regions = 0;
    void *heavy = nullptr, *heavy_g2 = nullptr;          // queues of oversized buckets for the CTA-wide fold
    cudaEvent_t ev_sorted = nullptr;                     // digit sort done (the G2 half may start on its own stream)
    uint32_t acc_threads_g1 = 0, acc_threads_g2 = 0;     // threads of one resident wave of the accumulate kernel (multiple of 128)
    int waves_alone = 1; bool alone = false;             // G1: waves to use when this MSM has the GPU to itself (set per run)
    void *buckets_g1 = nullptr, *buckets_g2 = nullptr;   // bucket pieces, slot = thread + bucket  [acc_threads + total + 1]
    void *out_g1 = nullptr, *out_g2 = nullptr;           // device partial sums  [(windows+1) * bpw]
    void *h_out_g1 = nullptr, *h_out_g2 = nullptr;       // pinned host copies
    cudaEvent_t ev_acc0 = nullptr, ev_acc1 = nullptr;    // around the G1 accumulate kernel (roofline measurement)
    // tails on team point operations (msm_team.cuh; ZKB200_TEAM=0 selects the one-thread-per-chain kernels of msm.cuh)
    int team = 7, team_now = 0;                          // bit 0: fold of the pieces, bit 1: fold of oversized buckets, bit 2: reduction; what the run in flight uses
    bool latency = true;                                 // this run has the GPU to itself (set per proof by prove_submit; stand-alone MSMs: yes)
    bool team_always = false;                            // ZKB200_TEAM_ALWAYS=1: also with other proofs in flight (measurement)
    int team_dbg = 0;                                    // ZKB200_TEAM_DBG: 1 host sums the CTA outputs, 2 one-thread scalar multiple, 4 one-thread point operations
    uint32_t seg_team = 0, bpw_team = 0;                 // buckets per chain in the weighted regions, CTAs (of 32 chains) per region
    void *team_counters = nullptr;                       // 2 x [regions + 1] (G1, G2) arrival counters of the device-side final sum, zero between launches
    void *final_g1 = nullptr, *final_g2 = nullptr;       // [regions + 1] one point per bucket region: what goes back to the host
    // batched-affine halving rounds in front of the XYZZ accumulation (msm.cuh); G1 only, 0 = off
    int affine_rounds = 0, rounds_now = 0;               // configured / used by the run in flight (a proof alone on the GPU skips them)
    bool affine_always = false;
    uint32_t aff_threads = 0;
    void *offsets_shifted = nullptr;                     // offsets >> affine_rounds
    void *aff_pts[4] = {nullptr, nullptr, nullptr, nullptr};   // output list of each round (affine points)
    void *aff_scratch = nullptr;                         // running denominator products, one Fq per pair of the first round
    float last_acc_ms() const;
    void init(uint32_t n, int c, uint32_t ones, bool g1, bool g2, bool expanded, int affine_rounds = 0);
    void release();
};
struct ScalarRef { const void *scalars; const uint32_t *map; uint32_t offset; int montgomery; };
// sort digits (count / scan / scatter) then accumulate+reduce for G1 and/or G2 bases; results land in plan.h_out_* after
// the stream is synchronised.
void msm_run(cudaStream_t st, MsmPlan &p, ScalarRef sc, const uint8_t *skip, const void *bases_g1, const void *bases_g2, cudaStream_t st_g2 = nullptr);
void *msm_expand_bases(const void *bases, uint32_t n, int c, bool g2);   // device table out[k*n+i] = 2^(c*k) * bases[i]
zkh::HG1 msm_finish_g1(const MsmPlan &p);      // host: add the partial sums (windowed layout: Horner over windows)
zkh::HG2 msm_finish_g2(const MsmPlan &p);

// ---- proving key resident on one GPU -------------------------------------------------------------------------------------
struct DeviceCsr { uint32_t *rowptr = nullptr, *col = nullptr, *coef = nullptr; uint32_t nnz = 0; };
// Everything ONE in-flight proof writes: assignment, QAP work vectors, MSM work areas, streams.  A proving key owns a few of
// these ("lanes") so that independent proofs overlap on the GPU: the latency-bound tails of one proof (bucket reduction, digit
// sort) run under the integer-bound kernels of the next, and host work (witness generation, proof assembly) overlaps both.
struct Lane {
    int index = 0;
    bool busy = false;
    void *w_can = nullptr, *w_mont = nullptr;            // (num_vars + 1) scalars: [1 | assignment]
    void *h_w_pinned = nullptr;
    void *w_lo = nullptr, *w_wide = nullptr;              // device: compact assignment staging
    void *h_wide_pinned = nullptr;
    void *d_seeds = nullptr, *h_seeds_pinned = nullptr;  // SHA-256 compression seeds of a GPU-expanded witness (prover.cu sha256_witness_kernel)
    void *bufA = nullptr, *bufB = nullptr, *bufC = nullptr, *tmp = nullptr;   // m Fr each
    uint32_t *sat_flag = nullptr; uint32_t *h_sat_flag = nullptr;
    MsmPlan mA, mB, mH, mL;
    cudaStream_t s_main = nullptr, s_a = nullptr, s_b = nullptr, s_l = nullptr, s_b2 = nullptr;
    cudaEvent_t ev_ja = nullptr, ev_jb = nullptr, ev_jl = nullptr, ev_jb2 = nullptr;     // joins of the side streams into s_main (ev_a .. ev_b2 are what the host polls)
    enum { SIG_A = 0, SIG_B, SIG_L, SIG_B2, SIG_DONE, NSIG };
    uint32_t *sig_dev = nullptr; volatile uint32_t *sig_host = nullptr; uint32_t sig_epoch = 0;   // completion signals: device counters, pinned host words, proofs submitted
    struct LaneGraph { cudaGraphExec_t exec = nullptr; int launches = 0; };
    LaneGraph graph[2];                                  // everything of a proof behind the upload, captured once: [0] other proofs in flight, [1] alone
    cudaStream_t s_sat = nullptr; cudaEvent_t ev_spmv = nullptr, ev_sat = nullptr;      // is_satisfied runs beside the first inverse transforms
    cudaEvent_t ev_w = nullptr, ev_a = nullptr, ev_b = nullptr, ev_l = nullptr, ev_b2 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_q0 = nullptr, ev_q1 = nullptr,
                ev_h0 = nullptr, ev_h1 = nullptr;
    uint64_t r[4] = {0, 0, 0, 0}, s[4] = {0, 0, 0, 0};   // zero-knowledge scalars of the proof in flight
    int launches = 0;                                    // kernels launched for the proof in flight
    unsigned long long h2d_bytes = 0;                    // bytes its upload copied to the device
    bool pending = false;                                // submitted, not yet collected
};
constexpr int MAX_LANES = 8;

struct DevicePk {
    int device = 0;
    uint64_t num_inputs = 0, num_vars = 0, num_constraints = 0;
    Domain *dom = nullptr;
    // bases (affine, Montgomery, AoS) and infinity flags
    void *A = nullptr, *B1 = nullptr, *B2 = nullptr, *H = nullptr, *L = nullptr;
    uint8_t *A_skip = nullptr, *B_skip = nullptr, *H_skip = nullptr, *L_skip = nullptr;
    uint32_t nA = 0, nB = 0, nH = 0, nL = 0;
    uint32_t *B_idx = nullptr, *L_idx = nullptr;
    zkh::HG1Affine alpha_g1, beta_g1, delta_g1;
    zkh::HG2Affine beta_g2, delta_g2;
    DeviceCsr a, b, c;
    void *coef_dict = nullptr; uint32_t ncoef = 0;
    // per-proof state
    Lane *lanes[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int nlanes = 0;
    void *sync = nullptr;                                 // mutex + condition variable guarding Lane::busy (opaque here)
    double load_seconds = 0, parse_seconds = 0, decompress_seconds = 0, expand_seconds = 0;
};

DevicePk *pk_load(const char *path, int device, std::string &err);     // lanes: env ZKB200_LANES (default 3)
// the two halves of pk_load, so that one parse of the key file can be made resident on several GPUs
zkpk::ParsedPk *pk_parse_file(const char *path, std::string &err, double *seconds);
void pk_parsed_free(zkpk::ParsedPk *P);
DevicePk *pk_from_parsed(const zkpk::ParsedPk &P, int device, std::string &err, double parse_seconds);
void pk_free(DevicePk *pk);
Lane *lane_acquire(DevicePk *pk);                       // blocks until a lane is free
Lane *lane_try(DevicePk *pk, int index);                // lane `index` if it is free, else nullptr
void lane_release(DevicePk *pk, Lane *ln);
Lane *lane_of_staging(DevicePk *pk, const void *p);     // the lane whose pinned staging buffer `p` points into, or nullptr

struct ProofPoints {
    zkh::HG1Affine A, C; zkh::HG2Affine B;
    zkh::HG1Affine At, Bt_h, Ht, Lt; zkh::HG2Affine Bt_g;   // the five MSM results (parity hooks)
    bool satisfied = true;
    bool want_parts = false;                            // also compute the five plain MSM values (costs four host scalar multiplications)
    float gpu_ms = 0, qap_ms = 0, msm_h_ms = 0, acc_h_ms = 0;   // CUDA-event timings of the last run (acc_h: H accumulate kernel)
    float a_done_ms = 0, b_done_ms = 0, l_done_ms = 0;          // when the A, B, L query MSMs (side streams) finished, from the start
    float host_tail_ms = 0;                             // host work left after the GPU finished (the rest overlaps the H-query MSM)
    int launches = 0;
};
// Two halves of one proof on a lane the caller holds.  submit: H2D copy of the assignment + every kernel, asynchronous.
//   assignment: num_vars canonical 32-byte LE scalars in HOST memory, or
//   lo/wide   : compact form (witness.hpp): lo[0..num_vars] = low 64 bits per variable (lo[0] = 1), wide = the few values above
//               64 bits; uploads 8 B per variable; lo may be the lane's own pinned staging (no staging copy), or
//   neither   : reuse the assignment already resident on the lane (bench "value" leg).
// collect: waits for the lane's streams, sums the partial points and assembles the proof on the host.
struct WideIn { uint32_t idx; uint32_t pad; uint64_t v[4]; };
// returns 0, or < 0 without touching the lane: -2 too many wide values, -3 wide index out of range, -4 r or s not below the group order
//   seeds     : (compact form only) `nseeds` records {uint32 base; uint32 w[16]; uint32 h[8]} of SHA-256 compression gadgets whose runs of
//               24 792 variables are not in lo (unset there) but expanded on the GPU; -5 if a run does not fit the assignment
int prove_submit(DevicePk *pk, Lane *ln, const uint8_t *assignment, const uint64_t *lo, const WideIn *wide, uint32_t nwide, const uint64_t r[4],
                 const uint64_t s[4], const void *seeds = nullptr, uint32_t nseeds = 0);
int lane_read_assignment(DevicePk *pk, Lane *ln, uint8_t *out);
int prove_collect(DevicePk *pk, Lane *ln, ProofPoints &out);
// synchronous conveniences: acquire a lane (or use the one that owns `lo`), submit, collect, release
int prove(DevicePk *pk, const uint8_t *assignment, const uint64_t r[4], const uint64_t s[4], ProofPoints &out);
int prove_compact(DevicePk *pk, const uint64_t *lo, const WideIn *wide, uint32_t nwide, const uint64_t r[4], const uint64_t s[4], ProofPoints &out,
                  const void *seeds = nullptr, uint32_t nseeds = 0);
uint64_t *compact_staging(Lane *ln);                    // pinned, (num_vars + 1) uint64
// QAP witness map only; writes (m+1)*32 bytes canonical to host `out_H`
int qap_witness_map(DevicePk *pk, const uint8_t *assignment, uint8_t *out_H, int *satisfied);

void device_init(int device);
void device_sync();                                 // cudaDeviceSynchronize of the current device, never while one of its streams is being captured
std::vector<int> devices_in_use();                  // every device device_init() has been called for
std::string proof_hex(const ProofPoints &p);           // mintcgo.cpp:112-187 layout
// measurement mode: the H-query MSM of the next proofs starts only after the A, B, L queries are done, so that the CUDA-event time of
// its kernels is that of the kernels alone (roofline); costs latency, never used otherwise
void set_isolate_h(bool on);
int launches_last_prove();                             // number of kernels launched by the last collected proof
void transfer_bytes_last_prove(unsigned long long out[2]);   // host-to-device, device-to-host bytes of the last proof this thread collected

} // namespace zkp
