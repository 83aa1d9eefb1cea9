// TEST INFRASTRUCTURE (oracle) -- not product code.
// C-ABI wrapper around the UNMODIFIED reference kernels of the prover hot path so
// tests and bench.py's CPU arm can call them through ctypes:
//   libfqfft evaluation domains  (FQ/evaluation_domain/get_evaluation_domain.tcc:33-52,
//                                 domains/basic_radix2_domain.tcc, domains/step_radix2_domain.tcc)
//   libff multi_exp              (FF/algebra/scalar_multiplication/multiexp.tcc:402-496)
//   libsnark kc_multi_exp        (LS/knowledge_commitment/kc_multiexp.tcc:21-89)
//   libsnark r1cs_to_qap_witness_map (LS/reductions/r1cs_to_qap/r1cs_to_qap.tcc:205-334)
// Built twice by oracle/Makefile: single-thread (shipped flags) and -DMULTICORE -fopenmp.
#include <boost/optional.hpp>
#include "libff/algebra/curves/alt_bn128/alt_bn128_pp.hpp"
#include "libff/algebra/scalar_multiplication/multiexp.hpp"
#include "libff/common/rng.hpp"
#include "libfqfft/evaluation_domain/get_evaluation_domain.hpp"
#include "libsnark/knowledge_commitment/kc_multiexp.hpp"
#include "libsnark/reductions/r1cs_to_qap/r1cs_to_qap.hpp"
#include "libsnark/relations/constraint_satisfaction_problems/r1cs/r1cs.hpp"
#ifdef MULTICORE
#include <omp.h>
#endif
#include "ref_io.h"

using namespace refio;

static void ensure_init() {
    static bool done = false;
    if (!done) { ppT::init_public_params(); done = true; }
}

extern "C" {

int ref_threads() {
#ifdef MULTICORE
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// returns the domain size m that get_evaluation_domain picks for min_size, and its kind
// (0 = basic_radix2, 1 = step_radix2, -1 = other/none).
long ref_domain_size(size_t min_size, int *kind) {
    ensure_init();
    try {
        auto d = libfqfft::get_evaluation_domain<FrT>(min_size);
        if (kind) {
            if (dynamic_cast<libfqfft::basic_radix2_domain<FrT> *>(d.get())) *kind = 0;
            else if (dynamic_cast<libfqfft::step_radix2_domain<FrT> *>(d.get())) *kind = 1;
            else *kind = -1;
        }
        return (long)d->m;
    } catch (...) { return -1; }
}

// op: 0 FFT, 1 iFFT, 2 cosetFFT(g=multiplicative_generator), 3 icosetFFT, 4 divide_by_Z_on_coset.
// data: m elements of 32 bytes (canonical LE), transformed in place.  Returns m or -1.
long ref_domain_op_timed(size_t min_size, int op, uint8_t *data, size_t n, double *seconds);
long ref_domain_op(size_t min_size, int op, uint8_t *data, size_t n) { return ref_domain_op_timed(min_size, op, data, n, nullptr); }
// same, *seconds (optional) = wall time of the transform alone (conversion of the bytes excluded)
long ref_domain_op_timed(size_t min_size, int op, uint8_t *data, size_t n, double *seconds) {
    ensure_init();
    try {
        auto d = libfqfft::get_evaluation_domain<FrT>(min_size);
        if (n != d->m) return -1;
        std::vector<FrT> a(n);
        for (size_t i = 0; i < n; i++) a[i] = get_fp<FrT>(data + 32 * i);
        const FrT g = FrT::multiplicative_generator;
        struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); const double t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
        switch (op) {
        case 0: d->FFT(a); break;
        case 1: d->iFFT(a); break;
        case 2: d->cosetFFT(a, g); break;
        case 3: d->icosetFFT(a, g); break;
        case 4: d->divide_by_Z_on_coset(a); break;
        default: return -1;
        }
        clock_gettime(CLOCK_MONOTONIC, &ts);
        if (seconds) *seconds = ts.tv_sec + 1e-9 * ts.tv_nsec - t0;
        for (size_t i = 0; i < n; i++) put_fp(data + 32 * i, a[i]);
        return (long)d->m;
    } catch (...) { return -1; }
}

// wall time of `reps` in-place forward FFTs of the domain over random-ish data already converted to Fr (conversion excluded)
double ref_fft_seconds(size_t min_size, int reps) {
    ensure_init();
    try {
        auto d = libfqfft::get_evaluation_domain<FrT>(min_size);
        std::vector<FrT> a(d->m);
        FrT x = FrT(12345), g = FrT(7);
        for (size_t i = 0; i < d->m; i++) { a[i] = x; x = x * g + FrT::one(); }
        struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); double t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
        for (int r = 0; r < reps; r++) d->FFT(a);
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return (ts.tv_sec + 1e-9 * ts.tv_nsec - t0) / reps;
    } catch (...) { return -1; }
}
// wall time of one multi_exp<G1, BDLO12> over n synthetic bases (running-sum points) and pseudo-random 254-bit scalars
double ref_msm_g1_seconds(size_t n, size_t chunks) {
    ensure_init();
    std::vector<G1T> b(n); std::vector<FrT> s(n);
    G1T d = FrT(31337) * G1T::one(), acc = d;
    for (size_t i = 0; i < n; i++) { b[i] = acc; acc = acc + d; }
    G1T::batch_to_special_all_non_zeros(b);
    FrT x = FrT(12345), g = FrT(7);
    for (size_t i = 0; i < n; i++) { s[i] = x; x = x * x + g; }
    if (chunks == 0) chunks = ref_threads();
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); double t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
    G1T r = libff::multi_exp<G1T, FrT, libff::multi_exp_method_BDLO12>(b.begin(), b.end(), s.begin(), s.end(), chunks);
    clock_gettime(CLOCK_MONOTONIC, &ts);
    volatile bool z = r.is_zero(); (void)z;
    return ts.tv_sec + 1e-9 * ts.tv_nsec - t0;
}

// domain element idx and vanishing polynomial at t (for the libfqfft-style property tests)
long ref_domain_element(size_t min_size, size_t idx, uint8_t *out) {
    ensure_init();
    try {
        auto d = libfqfft::get_evaluation_domain<FrT>(min_size);
        put_fp(out, d->get_domain_element(idx));
        return (long)d->m;
    } catch (...) { return -1; }
}

// method: 0 = multi_exp<BDLO12>, 1 = multi_exp_with_mixed_addition<BDLO12>.  seconds (optional) = wall time
// of the multi_exp call only (conversion excluded).
int ref_msm_g1(size_t n, const uint8_t *bases, const uint8_t *scalars, int method, size_t chunks, uint8_t *out, double *seconds) {
    ensure_init();
    std::vector<G1T> b(n); std::vector<FrT> s(n);
    for (size_t i = 0; i < n; i++) { b[i] = get_g1(bases + 64 * i); s[i] = get_fp<FrT>(scalars + 32 * i); }
    if (chunks == 0) chunks = ref_threads();
    double t0 = 0; struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
    G1T r = method == 0
        ? libff::multi_exp<G1T, FrT, libff::multi_exp_method_BDLO12>(b.begin(), b.end(), s.begin(), s.end(), chunks)
        : libff::multi_exp_with_mixed_addition<G1T, FrT, libff::multi_exp_method_BDLO12>(b.begin(), b.end(), s.begin(), s.end(), chunks);
    clock_gettime(CLOCK_MONOTONIC, &ts);
    if (seconds) *seconds = ts.tv_sec + 1e-9 * ts.tv_nsec - t0;
    put_g1(out, r);
    return 0;
}

int ref_msm_g2(size_t n, const uint8_t *bases, const uint8_t *scalars, int method, size_t chunks, uint8_t *out, double *seconds) {
    ensure_init();
    std::vector<G2T> b(n); std::vector<FrT> s(n);
    for (size_t i = 0; i < n; i++) { b[i] = get_g2(bases + 128 * i); s[i] = get_fp<FrT>(scalars + 32 * i); }
    if (chunks == 0) chunks = ref_threads();
    double t0 = 0; struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
    G2T r = method == 0
        ? libff::multi_exp<G2T, FrT, libff::multi_exp_method_BDLO12>(b.begin(), b.end(), s.begin(), s.end(), chunks)
        : libff::multi_exp_with_mixed_addition<G2T, FrT, libff::multi_exp_method_BDLO12>(b.begin(), b.end(), s.begin(), s.end(), chunks);
    clock_gettime(CLOCK_MONOTONIC, &ts);
    if (seconds) *seconds = ts.tv_sec + 1e-9 * ts.tv_nsec - t0;
    put_g2(out, r);
    return 0;
}

// Sparse (G2,G1) knowledge-commitment MSM exactly as the prover's B-query evaluation
// (r1cs_gg_ppzksnark.tcc:453-462): indices[k] sorted, values k -> (g2[k], g1[k]); scalars has domain_size entries.
int ref_kc_msm(size_t domain_size, size_t k, const uint64_t *indices, const uint8_t *g2s, const uint8_t *g1s,
               const uint8_t *scalars, size_t chunks, uint8_t *out_g2, uint8_t *out_g1, double *seconds) {
    ensure_init();
    libsnark::knowledge_commitment_vector<G2T, G1T> v;
    v.domain_size_ = domain_size;
    v.indices.resize(k); v.values.resize(k);
    for (size_t i = 0; i < k; i++) {
        v.indices[i] = indices[i];
        v.values[i] = libsnark::knowledge_commitment<G2T, G1T>(get_g2(g2s + 128 * i), get_g1(g1s + 64 * i));
    }
    std::vector<FrT> s(domain_size);
    for (size_t i = 0; i < domain_size; i++) s[i] = get_fp<FrT>(scalars + 32 * i);
    if (chunks == 0) chunks = ref_threads();
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); double t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
    auto r = libsnark::kc_multi_exp_with_mixed_addition<G2T, G1T, FrT, libff::multi_exp_method_BDLO12>(
        v, 0, domain_size, s.begin(), s.end(), chunks);
    clock_gettime(CLOCK_MONOTONIC, &ts);
    if (seconds) *seconds = ts.tv_sec + 1e-9 * ts.tv_nsec - t0;
    put_g2(out_g2, r.g); put_g1(out_g1, r.h);
    return 0;
}

// libff's deterministic field sampler (libff/common/rng.tcc:26-72), the scalar stream of the kernel sweep
int ref_sha512_rng(uint64_t first, size_t n, uint8_t *out) {
    ensure_init();
    for (size_t i = 0; i < n; i++) put_fp(out + 32 * i, libff::SHA512_rng<FrT>(first + i));
    return 0;
}

// k * P for G1/G2 (used to build test bases and to check the proof-assembly glue)
int ref_g1_mul(const uint8_t *p, const uint8_t *k, uint8_t *out) { ensure_init(); put_g1(out, get_fp<FrT>(k) * get_g1(p)); return 0; }
int ref_g2_mul(const uint8_t *p, const uint8_t *k, uint8_t *out) { ensure_init(); put_g2(out, get_fp<FrT>(k) * get_g2(p)); return 0; }
int ref_g1_add(const uint8_t *p, const uint8_t *q, uint8_t *out) { ensure_init(); put_g1(out, get_g1(p) + get_g1(q)); return 0; }
int ref_g2_add(const uint8_t *p, const uint8_t *q, uint8_t *out) { ensure_init(); put_g2(out, get_g2(p) + get_g2(q)); return 0; }
int ref_g1_gen(uint8_t *out) { ensure_init(); put_g1(out, G1T::one()); return 0; }
int ref_g2_gen(uint8_t *out) { ensure_init(); put_g2(out, G2T::one()); return 0; }

// n distinct bases  P_i = (start + i) * step * G  by a running sum (cheap: one Jacobian add per point),
// normalised to affine with the reference's batch_to_special_all_non_zeros.
int ref_g1_bases(size_t n, const uint8_t *step_scalar, uint8_t *out) {
    ensure_init();
    G1T d = get_fp<FrT>(step_scalar) * G1T::one();
    std::vector<G1T> v(n); G1T acc = d;
    for (size_t i = 0; i < n; i++) { v[i] = acc; acc = acc + d; }
    G1T::batch_to_special_all_non_zeros(v);
    for (size_t i = 0; i < n; i++) { put_fp(out + 64 * i, v[i].X); put_fp(out + 64 * i + 32, v[i].Y); }
    return 0;
}
int ref_g2_bases(size_t n, const uint8_t *step_scalar, uint8_t *out) {
    ensure_init();
    G2T d = get_fp<FrT>(step_scalar) * G2T::one();
    std::vector<G2T> v(n); G2T acc = d;
    for (size_t i = 0; i < n; i++) { v[i] = acc; acc = acc + d; }
    G2T::batch_to_special_all_non_zeros(v);
    for (size_t i = 0; i < n; i++) {
        put_fp(out + 128 * i, v[i].X.c0); put_fp(out + 128 * i + 32, v[i].X.c1);
        put_fp(out + 128 * i + 64, v[i].Y.c0); put_fp(out + 128 * i + 96, v[i].Y.c1);
    }
    return 0;
}

// Field ops (Montgomery-representation parity for the device field layer): out = a*b, a+b, a-b, a^-1 in Fr (which=0) or Fq (1)
int ref_field_op(int which, int op, size_t n, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    ensure_init();
    for (size_t i = 0; i < n; i++) {
        if (which == 0) {
            FrT x = get_fp<FrT>(a + 32 * i), y = b ? get_fp<FrT>(b + 32 * i) : FrT::one(), z;
            z = op == 0 ? x * y : op == 1 ? x + y : op == 2 ? x - y : (x.is_zero() ? x : x.inverse());
            put_fp(out + 32 * i, z);
        } else {
            FqT x = get_fp<FqT>(a + 32 * i), y = b ? get_fp<FqT>(b + 32 * i) : FqT::one(), z;
            z = op == 0 ? x * y : op == 1 ? x + y : op == 2 ? x - y : (x.is_zero() ? x : x.inverse());
            put_fp(out + 32 * i, z);
        }
    }
    return 0;
}
// raw Montgomery limbs of a canonical value (mont_repr = x * 2^256 mod p), to pin the device representation
int ref_to_mont(int which, size_t n, const uint8_t *a, uint8_t *out) {
    ensure_init();
    for (size_t i = 0; i < n; i++) {
        if (which == 0) { FrT x = get_fp<FrT>(a + 32 * i); memcpy(out + 32 * i, x.mont_repr.data, 32); }
        else { FqT x = get_fp<FqT>(a + 32 * i); memcpy(out + 32 * i, x.mont_repr.data, 32); }
    }
    return 0;
}

// r1cs_to_qap_witness_map on a caller-supplied R1CS in CSR form.
//   rowptr_X: num_constraints+1 entries; col_X: variable index (0 = ONE); coef_X: 32-byte canonical.
//   assignment: num_vars elements (primary first).  out_H: (m+1)*32 bytes.  Returns m or -1.
long ref_qap_witness_map(size_t num_inputs, size_t num_vars, size_t num_constraints,
                         const uint32_t *rowptr_a, const uint32_t *col_a, const uint8_t *coef_a,
                         const uint32_t *rowptr_b, const uint32_t *col_b, const uint8_t *coef_b,
                         const uint32_t *rowptr_c, const uint32_t *col_c, const uint8_t *coef_c,
                         const uint8_t *assignment, uint8_t *out_H, size_t out_cap, double *seconds) {
    ensure_init();
    try {
        libsnark::r1cs_constraint_system<FrT> cs;
        cs.primary_input_size = num_inputs;
        cs.auxiliary_input_size = num_vars - num_inputs;
        auto mk = [](const uint32_t *rp, const uint32_t *col, const uint8_t *coef, size_t i) {
            libsnark::linear_combination<FrT> lc;
            for (uint32_t k = rp[i]; k < rp[i + 1]; k++)
                lc.terms.emplace_back(libsnark::linear_term<FrT>(libsnark::variable<FrT>(col[k]), get_fp<FrT>(coef + 32 * (size_t)k)));
            return lc;
        };
        cs.constraints.reserve(num_constraints);
        for (size_t i = 0; i < num_constraints; i++)
            cs.constraints.emplace_back(libsnark::r1cs_constraint<FrT>(mk(rowptr_a, col_a, coef_a, i), mk(rowptr_b, col_b, coef_b, i), mk(rowptr_c, col_c, coef_c, i)));
        std::vector<FrT> prim(num_inputs), aux(num_vars - num_inputs);
        for (size_t i = 0; i < num_vars; i++) {
            FrT v = get_fp<FrT>(assignment + 32 * i);
            if (i < num_inputs) prim[i] = v; else aux[i - num_inputs] = v;
        }
        struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); double t0 = ts.tv_sec + 1e-9 * ts.tv_nsec;
        auto w = libsnark::r1cs_to_qap_witness_map(cs, prim, aux, FrT::zero(), FrT::zero(), FrT::zero());
        clock_gettime(CLOCK_MONOTONIC, &ts);
        if (seconds) *seconds = ts.tv_sec + 1e-9 * ts.tv_nsec - t0;
        size_t m = w.degree();
        if (out_cap < (m + 1) * 32) return -1;
        for (size_t i = 0; i <= m; i++) put_fp(out_H + 32 * i, w.coefficients_for_H[i]);
        return (long)m;
    } catch (...) { return -1; }
}

} // extern "C"
