/* zkb200.h -- C-ABI of the B200-native Groth16 prover for BlockMaze (libzkb200.so).
 *
 * Two layers, both plain C (pointers, sizes, integers -- no C++/CUDA/torch types):
 *
 *  (1) the reference's own cgo surface, re-exported verbatim so that go-ethereum/zktx/zktx.go links against this
 *      library in place of libzk_mint.so / libzk_send.so / libzk_deposit.so / libzk_redeem.so
 *      (reference headers: libsnark-vnt/src/mint/mintcgo.hpp:9-24, send/sendcgo.hpp:9-30, deposit/depositcgo.hpp:9-33,
 *      redeem/redeemcgo.hpp:9-24; Go call sites go-ethereum/zktx/zktx.go:122-550);
 *  (2) the zkb200_* entry points underneath: resident proving keys, the prover on a caller-supplied assignment,
 *      and the two hot kernels (evaluation-domain transforms, multi-scalar multiplication) for parity tests and the sweep.
 *
 * Wire formats of layer (2), all little-endian: field element = 32 bytes canonical (non-Montgomery); G1 affine = x y (64 B);
 * G2 affine = x.c0 x.c1 y.c0 y.c1 (128 B); all-zero bytes = point at infinity.
 * There is no CPU fallback: every compute entry point runs on the GPU selected with zkb200_init and aborts loudly on a
 * CUDA failure.
 */
#ifndef ZKB200_H
#define ZKB200_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------------------
 * (1) BlockMaze cgo surface.  256-bit arguments are "0x" + 64 hex chars, 160-bit ones "0x" + 40 (zktx.go:384-397);
 * returned strings are heap buffers owned by the caller exactly as in the reference (which never frees them).
 * Proof strings hold 512 lowercase hex chars + NUL: A.x A.y B.x.c1 B.x.c0 B.y.c1 B.y.c0 C.x C.y (mintcgo.cpp:176-187).
 * An unsatisfiable witness yields the "default proof" of the three group generators (mintcgo.cpp:207-211), which
 * starts with 63 '0' characters -- the condition go-ethereum tests (internal/ethapi/api.go:1486). */

/* libsnark-vnt/src/mint/mintcgo.hpp:9-24 (mintcgo.cpp:239-266, 268-321, 323-418) */
char *genCMT(uint64_t value, char *sn_string, char *r_string);
char *computePRF(char *sk_string, char *r_string);
char *genMintproof(uint64_t value, uint64_t value_old, char *sn_old_string, char *r_old_string, char *sn_string, char *r_string,
                   char *cmtA_old_string, char *cmtA_string, uint64_t value_s, char *sk_string);
bool verifyMintproof(char *data, char *cmtA_old_string, char *sn_old_string, char *cmtA_string, uint64_t value_s);

/* libsnark-vnt/src/send/sendcgo.hpp:9-30 (sendcgo.cpp:239-300, 302-368, 370-470) */
char *genCMTS(uint64_t value_s, char *pk_string, char *r_s_string, char *sn_old_string);
char *computeCRH(char *pk_string, char *r_string);
char *genSendproof(uint64_t value_A, char *r_s_string, char *sn_string, char *r_string, char *cmt_s_string, char *cmtA_string,
                   uint64_t value_s, char *pk_recv_string, uint64_t value_A_new, char *sn_A_new, char *r_A_new, char *cmt_A_new,
                   char *sk_string, char *pk_sender_string);
bool verifySendproof(char *data, char *cmtA_old_string, char *sn_old_string, char *cmtS_string, char *cmtA_new_string);

/* libsnark-vnt/src/deposit/depositcgo.hpp:9-33 (depositcgo.cpp:304-327, 329-448, 450-560) */
char *genRoot(char *cmtarray, int n);
char *genDepositproof(uint64_t value, uint64_t value_old, char *sn_old_string, char *r_old_string, char *sn_string, char *r_string,
                      char *sns_string, char *rs_string, char *cmtB_old_string, char *cmtB_string, uint64_t value_s, char *pk_string,
                      char *sn_A_oldstring, char *cmtS_string, char *cmtarray, int n, char *RT, char *sk_string);
bool verifyDepositproof(char *data, char *RT, char *pk, char *cmtb_old, char *snold, char *cmtb, char *sns);

/* libsnark-vnt/src/redeem/redeemcgo.hpp:9-24 (redeemcgo.cpp:269-322, 324-420) */
char *genRedeemproof(uint64_t value, uint64_t value_old, char *sn_old_string, char *r_old_string, char *sn_string, char *r_string,
                     char *cmtA_old_string, char *cmtA_string, uint64_t value_s, char *sk_string);
bool verifyRedeemproof(char *data, char *cmtA_old_string, char *sn_old_string, char *cmtA_string, uint64_t value_s);

/* ------------------------------------------------------------------------------------------------------------------
 * (2) zkb200 layer */

/* circuits, in the order used by every *_circuit argument */
enum { ZKB200_MINT = 0, ZKB200_SEND = 1, ZKB200_DEPOSIT = 2, ZKB200_REDEEM = 3 };

/* Devices.  Nothing pins a process to one GPU: a key handle runs on the device it is resident on, and the cgo surface of layer (1)
 * spreads its calls over the ACTIVE devices -- zkb200_set_devices() > env ZKB200_DEVICES ("all" or "0,2,3") > the device of an explicit
 * zkb200_init() (one process per GPU, as bench.py under torchrun) > env ZKB200_DEVICE > every visible device.  An unchanged geth process
 * therefore proves on all GPUs of the box (reference: one OpenMP knob, r1cs_gg_ppzksnark.tcc:429-433; zktx.go:383-404 stays as it is).
 *   zkb200_init           device of the layer-(2) calls that take no key handle (kernel entry points, benches).  0, or -1 if no GPU.
 *   zkb200_set_devices    active devices of layer (1); n = 0 selects every visible device.  Returns how many, or -1.  Resident keys are dropped.
 *   zkb200_active_devices copies the active device list, returns its length
 *   zkb200_last_device    device that proved this thread's last gen*proof
 *   zkb200_device_proofs  gen*proof calls this process has finished on `device` so far (what the scheduler dealt to it), or -1 */
int zkb200_init(int device);
int zkb200_ensure_device(void);
int zkb200_device_count(void);
int zkb200_current_device(void);
int zkb200_set_devices(const int *devices, int n);
int zkb200_active_devices(int *out, int cap);
int zkb200_last_device(void);
long zkb200_device_proofs(int device);
/* parity hook: copies the assignment (num_variables x 32 B canonical) that sits on the GPU behind this thread's last gen*proof call -- with
 * the SHA-256 gadget runs expanded on the device (default; ZKB200_GPU_WITNESS=0 has the host generator write them).  Returns the number of
 * variables or -1.  Only meaningful while no other caller has reused that lane. */
long zkb200_last_assignment(uint8_t *out, size_t cap_bytes);
/* Directory holding <circuit>pk.txt / <circuit>vk.txt.  Default: env ZKB200_KEY_DIR, else /usr/local/prfKey
 * (hard-coded in the reference: mintcgo.cpp:302,336). */
void zkb200_set_key_dir(const char *dir);
/* TEST HOOK -- pin the prover randomness: the next proofs draw (r, s) from this std::random_device-style 32-bit word stream exactly as
 * libff's Fr::random_element does (fp.tcc:695-721, bigint.tcc:167-179).  Refused (-1) unless the process runs with ZKB200_TEST_RNG=1;
 * the stream is never reused: when it is exhausted, or with n_words = 0, the prover is back on std::random_device. */
int zkb200_set_random_words(const uint32_t *words, size_t n_words);

/* Replaces r1cs_gg_ppzksnark_proving_key operator>> (r1cs_gg_ppzksnark.tcc:68-88): parse the file, decompress the points
 * on the GPU, keep everything resident.  Returns a handle or NULL (zkb200_last_error() says why). */
void *zkb200_pk_load(const char *path);                       /* on the zkb200_init device */
void *zkb200_pk_load_on(const char *path, int device);
/* one parse of the file, one resident copy per listed device (uploaded concurrently); out[i] belongs to devices[i].  0 or -1. */
int zkb200_pk_load_many(const char *path, const int *devices, int n, void **out);
int zkb200_pk_device(void *pk);
void zkb200_pk_free(void *pk);
/* info[0..7] = num_variables, num_inputs, num_constraints, domain m, domain kind (0 basic_radix2, 1 step_radix2),
 * nnz(A)+nnz(B)+nnz(C), distinct coefficients, B_query entries.  seconds[0..2] = total load, parse, GPU decompression + fixed-base table expansion. */
int zkb200_pk_info(void *pk, uint64_t info[8], double seconds[3]);
const char *zkb200_last_error(void);

/* The proof the reference returns for an unsatisfiable witness (r1cs_gg_ppzksnark_proof default constructor, r1cs_gg_ppzksnark.hpp:309-315):
 * 512 hex characters of (G1::one, G2::one, G1::one). */
const char *zkb200_default_proof(void);

/* Batch entry point over layer (1): n transactions, proved on every active device with all lanes busy (a pool of `threads` caller threads,
 * 0 = active devices x lanes, drains the list through gen*proof).  A transaction carries the arguments of its gen<Circuit>proof call:
 *   u[] = the uint64 arguments in order of appearance, s[] = the string arguments in order of appearance, n = genDepositproof's `n`.
 *   mint / redeem: u = value, value_old, value_s;        s = sn_old, r_old, sn, r, cmtA_old, cmtA, sk
 *   send         : u = value_A, value_s, value_A_new;    s = r_s, sn, r, cmt_s, cmtA, pk_recv, sn_A_new, r_A_new, cmt_A_new, sk, pk_sender
 *   deposit      : u = value, value_old, value_s;        s = sn_old, r_old, sn, r, sns, rs, cmtB_old, cmtB, pk, sn_A_old, cmtS, cmtarray, RT, sk
 * proofs: n x 513 bytes (512 hex + NUL each).  Returns the number of default proofs (unsatisfiable transactions), or -1 on bad arguments. */
typedef struct zkb200_tx { int circuit; int n; uint64_t u[3]; const char *s[14]; } zkb200_tx;
int zkb200_prove_batch(size_t n, const zkb200_tx *txs, char *proofs, int threads);
/* Batch verification (callers core/tx_pool.go:619-641, core/state_processor.go:113-158 verify every zk transaction of a
 * block one by one with r1cs_gg_ppzksnark_verifier_strong_IC, r1cs_gg_ppzksnark.tcc:524-623).  One item = the arguments of one
 * verify<Circuit>proof call: s[] = its string arguments after the proof, in order; value_s for mint / redeem.  Circuits may be mixed.
 * A random linear combination checks all proofs with one final exponentiation; ok[i] receives the verdict of item i, identical to
 * verify<Circuit>proof's (if the combined check fails, or a proof's B lies outside the order-r subgroup, the proofs concerned are checked
 * one by one).  threads: host threads to use, 0 = all.  Returns the number of valid proofs, or -1 on bad arguments / unreadable key. */
typedef struct zkb200_vtx { int circuit; uint64_t value_s; const char *proof; const char *s[6]; } zkb200_vtx;
int zkb200_verify_batch(size_t n, const zkb200_vtx *items, uint8_t *ok, int threads);
/* test hook: the two G2 membership tests of the batch verifier on one point of the twist curve (x.c0 x.c1 y.c0 y.c1, canonical):
 * bit 0 = psi(Q) == [6z^2]Q (the fast test it uses), bit 1 = [r]Q == O (the definition) */
int zkb200_g2_subgroup_check(const uint8_t point[128]);
/* the device-dispatch policy on its own (host only; CPU unit tests): least proofs in flight, ties round-robin */
void *zkb200_sched_create(int n_devices);
int zkb200_sched_pick(void *sched);
void zkb200_sched_done(void *sched, int slot);
int zkb200_sched_inflight(void *sched, int slot);
void zkb200_sched_free(void *sched);

/* Replaces r1cs_gg_ppzksnark_prover (r1cs_gg_ppzksnark.tcc:390-506) for a caller-supplied full assignment
 * (primary || auxiliary, num_variables x 32 B) and explicit r, s.  assignment == NULL re-proves the assignment already
 * resident on the GPU.  proof_hex: 513 bytes.  parts (optional, 384 B): the five MSM results At | Bt.g | Bt.h | Ht | Lt.
 * timings_ms (optional, 8 floats): GPU total, QAP witness map, H MSM, host work left after the GPU finished (most of the proof assembly
 * overlaps the H MSM), H-MSM bucket-accumulate kernel, and the times at
 * which the A, B and L query MSMs (side streams) were done, counted from the start (CUDA events).
 * Returns 0 = proof, 1 = constraint system not satisfied (proof_hex = default proof), <0 = error (-2 more than 61 wide values,
 * -3 a wide value's index outside [1, num_variables], -4 r or s not below the group order): malformed input never yields a proof. */
int zkb200_prove(void *pk, const uint8_t *assignment, const uint8_t r[32], const uint8_t s[32], char *proof_hex, uint8_t *parts,
                 float *timings_ms);
/* Same prover fed with the COMPACT assignment the native witness generators produce: lo[0..num_variables] = low 64 bits of every
 * variable (lo[0] = 1, the constant ONE), wide = nwide records {uint32 idx; uint32 pad; uint64 v[4]} for the few values above 64 bits.
 * 8 instead of 32 bytes per variable cross PCIe.  lo may be zkb200_lane_staging() of a lane the caller holds (pinned: no staging copy;
 * the proof then runs on that lane).
 * seeds (optional): nseeds records {uint32 base; uint32 w[16]; uint32 h[8]}, one per sha256_compression_function_gadget instance whose run of
 * 24 792 internal variables (starting at variable `base`) the GPU fills in from the 16 message words w and the incoming chaining value h
 * (witness generation on the GPU, SURVEY.md 8f rank 3: these runs are 97 % of a BlockMaze assignment); lo need not be set there.
 * Returns -5 if a run does not fit the assignment or nseeds > 32. */
int zkb200_prove_compact(void *pk, const uint64_t *lo, const void *wide, size_t nwide, const void *seeds, size_t nseeds, const uint8_t r[32],
                         const uint8_t s[32], char *proof_hex, float *timings_ms);

/* Proofs in flight.  A resident proving key owns zkb200_pk_lanes(pk) "lanes" (env ZKB200_LANES, default 3): private copies of every
 * buffer one proof writes, with their own CUDA streams.  zkb200_prove / zkb200_prove_compact / gen*proof are thread-safe and take a
 * free lane each, so concurrent callers (goroutines in geth) overlap on the GPU.  A single-threaded caller gets the same overlap by
 * splitting a proof in two: submit enqueues the copy and every kernel and returns, collect waits and assembles the proof.
 *   zkb200_lane_acquire  returns a free lane index (blocks while all are busy); zkb200_lane_release gives it back
 *   zkb200_lane_staging  pinned buffer of num_variables + 1 uint64 for the compact assignment of that lane
 *   zkb200_prove_submit  assignment (num_variables x 32 B, host) or NULL = the assignment the lane already holds
 *   zkb200_prove_submit_compact  lo / wide / seeds as in zkb200_prove_compact
 *   zkb200_prove_collect same outputs and return value as zkb200_prove (timings_ms[3], host finish, counts from the call) */
int zkb200_pk_lanes(void *pk);
int zkb200_lane_acquire(void *pk);
void zkb200_lane_release(void *pk, int lane);
uint64_t *zkb200_lane_staging(void *pk, int lane);
int zkb200_prove_submit(void *pk, int lane, const uint8_t *assignment, const uint8_t r[32], const uint8_t s[32]);
int zkb200_prove_submit_compact(void *pk, int lane, const uint64_t *lo, const void *wide, size_t nwide, const void *seeds, size_t nseeds,
                                const uint8_t r[32], const uint8_t s[32]);
/* parity hook: the canonical assignment (num_variables x 32 B) lane `lane` holds, i.e. what its last proof was made for */
int zkb200_lane_read_assignment(void *pk, int lane, uint8_t *out);
int zkb200_prove_collect(void *pk, int lane, char *proof_hex, uint8_t *parts, float *timings_ms);
/* Replaces r1cs_to_qap_witness_map (r1cs_to_qap.tcc:205-334): out_H receives (m+1) x 32 B coefficients_for_H. */
int zkb200_qap_witness_map(void *pk, const uint8_t *assignment, uint8_t *out_H, int *satisfied);
/* milliseconds of the last gen*proof call: host witness generation, zkb200_prove total, of which GPU (CUDA events), host finish */
void zkb200_last_breakdown_ms(double out[4]);
/* kernels launched for the last collected proof */
int zkb200_last_launches(void);
/* bytes the last proof this thread made (gen*proof, zkb200_prove*) moved over PCIe: out[0] host to device (the assignment upload -- 60 KB for a
 * send proof with the SHA-256 runs expanded on the GPU), out[1] device to host (partial sums of the four queries, the is_satisfied flag) */
void zkb200_last_transfer_bytes(unsigned long long out[2]);
/* measurement mode (bench.py roofline): when on, a proof's H-query MSM starts only after its A, B, L queries are done, so the CUDA-event
 * time of the H kernels (timings_ms[2], [4]) is that of the kernels running alone.  Costs latency; off by default. */
void zkb200_set_isolate_h(int on);

/* Witness handling: the FULL variable assignment (primary || auxiliary, 32 B canonical each) that the reference obtains by running
 * its gadgetlib1 circuit (<circuit>_gadget::generate_r1cs_witness, SRC/<c>/circuit/gadget.tcc), computed natively on the host.
 * Arguments are exactly those of gen<Circuit>proof (redeem != 0 selects the redeem circuit, which shares mint's signature).
 * Returns the number of variables written, -1 if cap (in elements) is too small, -2 if cmtS is not among the deposit leaves. */
/* parity hook: on != 0 makes the three functions below (on this thread) run the generator the way gen*proof does by default -- the runs of
 * the SHA-256 compression gadgets left to a separate expansion from 100-byte seeds (on the GPU in gen*proof; here its host copy,
 * csrc/witness_sha.hpp) -- so that both paths can be compared variable for variable */
void zkb200_witness_defer(int on);
long zkb200_witness_mint(uint64_t value, uint64_t value_old, const char *sn_old, const char *r_old, const char *sn, const char *r,
                         const char *cmtA_old, const char *cmtA, uint64_t value_s, const char *sk, int redeem, uint8_t *out, size_t cap);
long zkb200_witness_send(uint64_t value_A, const char *r_s, const char *sn, const char *r, const char *cmt_s, const char *cmtA, uint64_t value_s,
                         const char *pk_recv, uint64_t value_A_new, const char *sn_A_new, const char *r_A_new, const char *cmt_A_new,
                         const char *sk, const char *pk_sender, uint8_t *out, size_t cap);
long zkb200_witness_deposit(uint64_t value, uint64_t value_old, const char *sn_old, const char *r_old, const char *sn, const char *r,
                            const char *sns, const char *rs, const char *cmtB_old, const char *cmtB, uint64_t value_s, const char *pk,
                            const char *sn_A_old, const char *cmtS, const char *cmtarray, int n, const char *RT, const char *sk, uint8_t *out,
                            size_t cap);

/* Replaces r1cs_gg_ppzksnark_generator + the *_key tools (r1cs_gg_ppzksnark.tcc:211-388, SRC/<c>/getpvk.cpp) for the constraint system
 * embedded in an existing proving-key file: draws t, alpha, beta, gamma, delta and the two generator scalars from `words` (same
 * consumption as Fr::random_element; n_words = 0: OS entropy), evaluates the QAP at t on the host and does the ~2.3 M fixed-base
 * scalar multiplications on the GPU, then writes pk / vk files in the reference's format.  seconds[0..2] = host QAP evaluation,
 * GPU fixed-base phase, file encoding.  Returns 0 on success. */
int zkb200_keygen(const char *cs_source_pk_path, const uint32_t *words, size_t n_words, const char *out_pk_path, const char *out_vk_path,
                  double seconds[3]);

/* Replaces libfqfft::get_evaluation_domain(min_size) + domain->{FFT,iFFT,cosetFFT,icosetFFT,divide_by_Z_on_coset}
 * (get_evaluation_domain.tcc:33-52, basic_radix2_domain.tcc:26-112, step_radix2_domain.tcc:21-248).
 * op: 0 FFT, 1 iFFT, 2 cosetFFT (g = 5), 3 icosetFFT, 4 divide_by_Z_on_coset.  data: n = domain size elements, in place.
 * Returns the domain size m (call with data == NULL to query it; *kind = 0 basic, 1 step), or -1. */
long zkb200_domain_op(size_t min_size, int op, uint8_t *data, size_t n, int *kind);

/* Replaces libff::multi_exp<G, Fr, multi_exp_method_BDLO12> and multi_exp_with_mixed_addition (multiexp.tcc:402-496).
 * window_bits = 0 picks a default. */
int zkb200_msm_g1(size_t n, const uint8_t *bases, const uint8_t *scalars, int window_bits, uint8_t out[64]);
int zkb200_msm_g2(size_t n, const uint8_t *bases, const uint8_t *scalars, int window_bits, uint8_t out[128]);

/* Device field layer check: out = a (op) b on raw MONTGOMERY representatives, computed by a CUDA kernel.
 * field: 0 Fr, 1 Fq.  op: 0 mul, 1 add, 2 sub, 3 sqr(a), 4 to_mont(a), 5 from_mont(a), 6 inverse(a). */
int zkb200_field_op(int field, int op, size_t n, const uint8_t *a, const uint8_t *b, uint8_t *out);

/* Device-resident benchmarks (inputs generated / kept in HBM, timed with CUDA events on the launching stream).
 * zkb200_bench_ntt: `iters` forward size-2^logn transforms over a buffer of `batch` independent vectors; returns ms per transform.
 * zkb200_bench_msm: dense 254-bit MSM over n synthetic bases and scalars (zkb200_synth_*; group: 1 = G1, 2 = G2); returns ms per MSM.  window_bits > 0:
 * windowed layout (bases as given, Horner on the host); window_bits < 0: fixed-base layout with |window_bits| bits (the table
 * 2^(c*k)*P is built once outside the timed region, as for a resident proving key); 0: default windowed. */
float zkb200_bench_ntt(int logn, int batch, int iters);
float zkb200_bench_msm(int group, size_t n, int window_bits, int iters);
/* The same synthetic problem restricted to points [first, first+n): bases and scalars are functions of the GLOBAL index, so a single
 * large MSM splits by point range over several GPUs (one partial point per GPU, added on the host).  out_point
 * (64 B G1 / 128 B G2, may be NULL) receives the partial sum. */
float zkb200_bench_msm_slice(int group, size_t first, size_t n, int window_bits, int iters, uint8_t *out_point);
/* Host-side sum of n affine G1 (64 B each as in zkb200_msm_g1) / G2 (128 B) points, all-zero = infinity: the "one partial point per GPU
 * summed on the host" step of an MSM split by point range. */
int zkb200_g1_sum(size_t n, const uint8_t *points, uint8_t out[64]);
int zkb200_g2_sum(size_t n, const uint8_t *points, uint8_t out[128]);
/* The synthetic inputs of the sweep as canonical bytes, so that the host legs (libff multi_exp, libfqfft FFT) see IDENTICAL inputs:
 * scalar i = libff SHA512_rng<Fr>(i) (libff/common/rng.tcc:26-72); base i = SHA512_rng<Fr>(2^32 + i) * generator (group 1 = G1, 2 = G2). */
int zkb200_synth_scalars(size_t first, size_t n, uint8_t *out);
int zkb200_synth_bases(int group, size_t first, size_t n, uint8_t *out);
/* bench hygiene, over every device in use: overwrite a 256 MB scratch buffer (2x L2) and synchronise; plain cudaDeviceSynchronize */
void zkb200_flush_l2(void);
void zkb200_device_sync(void);
/* device-clock stopwatch for a region spread over many streams: synchronises the device, then records a CUDA event.
 * stop = 0 starts, stop = 1 returns the milliseconds elapsed on the GPU since the start */
float zkb200_device_timer(int stop);
/* integer-multiply peaks of the GPU in 1e12 ops/s (the MSM roofline denominators): mode 0 = 32x32->32 multiply-add (IMAD),
 * mode 1 = 32x32->64 multiply-add in carry chains (IMAD.WIDE.U32.X, the instruction the field multiplication is made of),
 * mode 2 = whole 254-bit Montgomery multiplications */
float zkb200_bench_imad_peak(int mode);

#ifdef __cplusplus
}
#endif
#endif
