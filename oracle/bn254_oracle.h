/*
 * oracle/bn254_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, gcc, no dependencies) of the reference's BN254 G1
 * variable-base MSM path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (plonkish_b200/, include/) never links, imports or calls it.
 *
 * Parity status: "parity unpinned" against reference *binaries* — the
 * reference is Rust (no cargo/rustc here) and holds no golden vectors for
 * msm.rs.  The oracle is pinned instead by (1) an independent
 * Python big-int implementation (oracle/bigint_ref.py), (2) public BN254 /
 * EIP-196 known answers, (3) known-discrete-log identities; see
 * tests/test_oracle.py and tests/golden/.
 *
 * What is restated, with the reference lines it follows
 * (paths relative to /root/reference/plonkish_backend/src):
 *   util/arithmetic/msm.rs:8-14     window_size
 *   util/arithmetic/msm.rs:33-48    windowed_scalar
 *   util/arithmetic/msm.rs:84-115   variable_base_msm (T-way point chunking)
 *   util/arithmetic/msm.rs:117-181  variable_base_msm_serial + CurveAcc
 *   util/parallel.rs:9-25           parallelize_iter (one task per chunk)
 *   util/transcript.rs:216-229      write_commitment byte encoding
 *   util/arithmetic/msm.rs:16-31,50-81  window_table, fixed_base_msm
 *   pcs/multilinear.rs:72-107, 203-213  quotients, g_prime merge
 *   pcs/multilinear/kzg.rs:174-208      eq tables of the SRS
 *   piop/sum_check/classic.rs:90-141, classic/eval.rs:101-131, poly/multilinear.rs:179-189   sum-check rounds
 * The field / curve arithmetic the reference gets from the un-vendored
 * third-party crate halo2_curves 0.3.3 (plonkish_backend/Cargo.toml:7, patched
 * at Cargo.toml:9-11, no lockfile) is restated from the published BN254
 * (alt_bn128) definition: y^2 = x^3 + 3 over Fq, G = (1, 2), 4x64-bit
 * little-endian Montgomery limbs with R = 2^256, affine identity = (0, 0),
 * Jacobian identity z = 0.
 */
#ifndef PLONKISH_ORACLE_BN254_H
#define PLONKISH_ORACLE_BN254_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } ofe_t;            /* Fq or Fr element, Montgomery form */
typedef struct { ofe_t x, y; } og1_affine_t;        /* 64 B; (0,0) = identity            */
typedef struct { ofe_t x, y, z; } og1_jac_t;        /* 96 B Jacobian; z = 0 = identity   */

/* ---- field helpers (which = 0 -> Fq, 1 -> Fr) ---- */
void oracle_fe_mul(int which, const ofe_t *a, const ofe_t *b, ofe_t *out);
void oracle_fe_add(int which, const ofe_t *a, const ofe_t *b, ofe_t *out);
void oracle_fe_sub(int which, const ofe_t *a, const ofe_t *b, ofe_t *out);
void oracle_fe_inv(int which, const ofe_t *a, ofe_t *out);          /* 0 -> 0 */
void oracle_fe_from_canonical(int which, const uint64_t c[4], ofe_t *out);
void oracle_fe_to_canonical(int which, const ofe_t *a, uint64_t c[4]); /* == to_repr (LE limbs) */

/* ---- curve helpers ---- */
void oracle_g1_generator(og1_affine_t *out);
int  oracle_g1_is_on_curve(const og1_affine_t *p);                 /* identity counts as on-curve */
void oracle_g1_add(const og1_jac_t *a, const og1_jac_t *b, og1_jac_t *out);
void oracle_g1_add_mixed(const og1_jac_t *a, const og1_affine_t *b, og1_jac_t *out);
void oracle_g1_double(const og1_jac_t *a, og1_jac_t *out);
void oracle_g1_to_affine(const og1_jac_t *a, og1_affine_t *out);
void oracle_g1_from_affine(const og1_affine_t *a, og1_jac_t *out);
/* k = canonical (non-Montgomery) 256-bit little-endian limbs */
void oracle_g1_scalar_mul(const og1_affine_t *base, const uint64_t k[4], og1_jac_t *out);

/* transcript.rs:216-229 — x||y, each 32-byte big-endian canonical. Returns 0 on
 * success, -1 for the identity (the reference's coordinates().unwrap() panics). */
int oracle_g1_transcript_bytes(const og1_affine_t *p, uint8_t out[64]);

/* ---- the hot path ---- */
size_t oracle_window_size(size_t num_scalars);                          /* msm.rs:8-14  */
size_t oracle_windowed_scalar(size_t window_size, size_t window_mask,
                              size_t idx, const uint8_t repr[32]);      /* msm.rs:33-48 */
/* msm.rs:117-181; accumulates into *result exactly as the reference does. */
void oracle_variable_base_msm_serial(const ofe_t *scalars, const og1_affine_t *bases,
                                     size_t n, og1_jac_t *result);
/* msm.rs:84-115 with num_threads() == T.  n == 0 returns the identity (the
 * reference panics at msm.rs:154; documented deviation). */
void oracle_variable_base_msm(const ofe_t *scalars, const og1_affine_t *bases, size_t n,
                              int num_threads, og1_jac_t *out);
/* Independent check: plain double-and-add per point, summed. Small n only. */
void oracle_msm_naive(const ofe_t *scalars, const og1_affine_t *bases, size_t n, og1_jac_t *out);

/* ---- synthetic inputs with a known answer ----
 * bases[i] = (a + i*d) * G  (a, d canonical 256-bit LE limbs, small in practice),
 * so sum_i s_i * bases[i] = ((a * sum s_i + d * sum i*s_i) mod r) * G.           */
void oracle_known_dlog_bases(const uint64_t a[4], const uint64_t d[4], size_t n,
                             int num_threads, og1_affine_t *out);
void oracle_known_dlog_answer(const uint64_t a[4], const uint64_t d[4],
                              const ofe_t *scalars, size_t n, og1_affine_t *out);

/* ---- the callers either side of the MSM ---- */
/* pcs/multilinear.rs:72-107: quotients[] has 2^num_vars entries, q_i (2^i values) at offset 2^i. */
void oracle_quotients(const ofe_t *evals, const ofe_t *point, size_t num_vars, ofe_t *quotients, ofe_t *eval_out);
/* pcs/multilinear.rs:203-213 (g_prime): out[j] = sum_i coeffs[i] * polys[i][j]. */
void oracle_fr_linear_combination(const ofe_t *const *polys, const ofe_t *coeffs, size_t count, size_t n, ofe_t *out);
/* pcs/multilinear/kzg.rs:174-193: the eq tables as scalars, slice k (2^k values) at offset 2^k - 1. */
void oracle_kzg_eq_scalars(const ofe_t *ss, size_t num_vars, ofe_t *out);
/* msm.rs:16-31 (window_table), :50-81 (fixed_base_msm) and batch_normalize (kzg.rs:204-207):
 * out[i] = scalars[i] * base, affine. */
void oracle_fixed_base_msm(const og1_affine_t *base, size_t window_size, const ofe_t *scalars, size_t n, int num_threads,
                           og1_affine_t *out);

/* ---- sum check ---- */
/* piop/sum_check/classic/eval.rs:101-131 on explicit tables of 2*size evaluations: out[x-1] = sum_b expr(X = x),
 * x = 1..degree, expr = sum_t coeffs[t] * prod polys[term_polys[offsets[t]..offsets[t+1])] (* polys[common]). */
void oracle_sumcheck_round(const ofe_t *const *polys, size_t num_polys, size_t size, const ofe_t *coeffs,
                           const uint32_t *offsets, const uint32_t *term_polys, size_t num_terms, int common,
                           size_t degree, ofe_t *out);
/* MultilinearPolynomial::fix_var (poly/multilinear.rs:179-189, 599-618): n evaluations in, n/2 out. */
void oracle_fix_var(const ofe_t *evals, size_t n, const ofe_t *x, ofe_t *out);
/* The same two loops cut over num_threads pthreads (identical results: field addition is exact). */
void oracle_sumcheck_round_mt(const ofe_t *const *polys, size_t num_polys, size_t size, const ofe_t *coeffs,
                              const uint32_t *offsets, const uint32_t *term_polys, size_t num_terms, int common,
                              size_t degree, int num_threads, ofe_t *out);
void oracle_fr_vec_op(int op, const ofe_t *a, const ofe_t *b, size_t n, int num_threads, ofe_t *out);
void oracle_fr_affine(const ofe_t *const *polys, const uint32_t *const *rows, const ofe_t *coeffs, size_t count, const ofe_t *constant,
                      const ofe_t *id_coeff, size_t n, int num_threads, ofe_t *out);
void oracle_fix_var_mt(const ofe_t *evals, size_t n, const ofe_t *x, int num_threads, ofe_t *out);

/* Keccak256 (original Keccak padding) as used by Keccak256Transcript (util/transcript.rs:100-131, util/hash.rs:5-8). */
void oracle_keccak256(const uint8_t *data, size_t len, uint8_t out[32]);

/* div_rem by (X - z) (poly/univariate.rs:144-168 as UnivariateKzg::open calls it, pcs/univariate/kzg.rs:281-282). */
void oracle_fr_div_linear(const ofe_t *coeffs, size_t n, const ofe_t *z, ofe_t *q, ofe_t *rem);

/* BooleanHypercube::iter (util/arithmetic/bh.rs:118-125) and permutation_z_polys (backend/hyperplonk/prover.rs:252-345). */
void oracle_bh_iter(size_t num_vars, uint32_t *out);
void oracle_permutation_z_polys(size_t num_chunks, const ofe_t *const *values, const ofe_t *const *sigmas, size_t count, size_t num_vars,
                                const ofe_t *beta, const ofe_t *gamma, ofe_t *out);
void oracle_permutation_z_polys_mt(size_t num_chunks, const ofe_t *const *values, const ofe_t *const *sigmas, size_t count, size_t num_vars,
                                   const ofe_t *beta, const ofe_t *gamma, int num_threads, ofe_t *out);

/* Array forms of oracle_fe_from_canonical / oracle_fe_to_canonical. */
void oracle_fe_from_canonical_n(int which, const uint64_t *in, size_t n, uint64_t *out);
void oracle_fe_to_canonical_n(int which, const uint64_t *in, size_t n, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
