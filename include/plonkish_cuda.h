/*
 * plonkish_cuda.h — C ABI of libplonkish_cuda.so: BN254 G1 variable-base MSM on
 * NVIDIA B200 (sm_100a).  This is the drop-in boundary for
 *
 *     plonkish_backend::util::arithmetic::variable_base_msm(scalars, bases)
 *     /root/reference/plonkish_backend/src/util/arithmetic/msm.rs:84-115
 *
 * The reference has no FFI of its own (100 % Rust); these are the
 * entry points the `plonkish_cuda` Rust crate binds (bindings/plonkish_cuda/src/lib.rs,
 * INTEGRATION.md) from inside that function when C == bn256::G1Affine.
 *
 * Data crossing the boundary (all little-endian, exactly the in-memory layout
 * of halo2_curves 0.3.3 [ext] types, plonkish_backend/Cargo.toml:7):
 *   scalar  32 B  bn256::Fr       4 x u64 limbs, Montgomery form (a * 2^256 mod r)
 *   base    64 B  bn256::G1Affine x || y, each a Montgomery Fq; (0, 0) = identity
 *   result  64 B  bn256::G1Affine same encoding; the identity comes back as (0, 0)
 * The reference returns a projective `C::Curve` that every hot-path caller
 * normalises at once (`.into()` at pcs/multilinear/kzg.rs:255,271,292 and
 * pcs/univariate/kzg.rs:28; `.to_affine()` at pcs.rs:175); the library returns
 * that affine value, bit-exact.
 *
 * Conventions: every function returns 0 on success or a negative PLONKISH_CUDA_E_*
 * code; plonkish_cuda_last_error() gives the message for the calling thread.
 * There is no CPU fallback: without a usable CUDA device every compute call fails.
 * Calls are blocking unless they take a stream, and thread-safe (the reference
 * calls variable_base_msm from rayon workers, pcs/multilinear/hyrax.rs:176-180).
 * n == 0 returns the identity (the reference panics at msm.rs:154; documented
 * deviation).  Mismatched lengths cannot be expressed: one `n` covers both arrays
 * (the reference asserts equality at msm.rs:90).
 */
#ifndef PLONKISH_CUDA_H
#define PLONKISH_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLONKISH_CUDA_OK 0
#define PLONKISH_CUDA_E_INVALID (-1)   /* bad argument (null pointer, unknown handle, device out of range) */
#define PLONKISH_CUDA_E_CUDA (-2)      /* a CUDA runtime call failed; see last_error */
#define PLONKISH_CUDA_E_NO_DEVICE (-3) /* no CUDA device / library not initialised */
#define PLONKISH_CUDA_E_COMM (-4)      /* multi-GPU exchange failed */

#define PLONKISH_CUDA_SCALAR_BYTES 32
#define PLONKISH_CUDA_AFFINE_BYTES 64
#define PLONKISH_CUDA_XYZZ_BYTES 128 /* projective partial: X, Y, ZZ, ZZZ (x = X/ZZ, y = Y/ZZZ); ZZ = 0 is the identity */

/* ---- lifetime ------------------------------------------------------------------
 * Creates one context (stream, scratch arena, base cache) per device for the first
 * n_devices CUDA devices; n_devices <= 0 means all visible devices.  Idempotent.
 * Replaces nothing in the reference (rayon's pool is implicit, util/parallel.rs:1-7). */
int plonkish_cuda_init(int n_devices);
int plonkish_cuda_device_count(void); /* devices made available by init (contexts are created on first use) */
void plonkish_cuda_shutdown(void);
const char *plonkish_cuda_last_error(void);

/* ---- resident bases --------------------------------------------------------------
 * Uploads n affine bases to `device` and returns a handle.  The reference re-reads
 * its SRS from host memory on every call (ProverParam.eqs, pcs/multilinear/kzg.rs:55-77;
 * powers_of_s_g1, pcs/univariate/kzg.rs:24-30); the SRS is static per ProverParam, so
 * a ProverParam wrapper registers each slice once and passes the handle afterwards.  A later
 * MSM may use any prefix n' <= n of a registered slice (a prefix shorter than 1/16 of the
 * slice runs on the plain bases, row 0 of the table).  Handles are reference counted:
 * release while another thread's call still uses the slice frees the memory when that call
 * is done. */
int plonkish_cuda_bases_register(int device, const void *bases_affine64, size_t n, uint64_t *handle);
int plonkish_cuda_bases_release(uint64_t handle);

/* Borrowed slices.  Inside variable_base_msm (msm.rs:84-87) the shim only sees `&[G1Affine]`: it
 * cannot tell a ProverParam's static SRS slice from a temporary whose address the allocator will
 * reuse (ipa.rs g_lo/g_hi, hyrax.rs rows, a second setup in one process).  bases_cached looks the
 * address up in a cache inside the library and trusts a hit only if the content fingerprint taken
 * at registration (first 8 points, the last one, 24 pseudo-random positions; for a prefix also the
 * prefix's last point against the device copy) matches what is at that address now; otherwise the
 * stale entry is released and the slice registered afresh.  A longer slice at a cached address
 * replaces the entry; a shorter one uses its prefix (&powers_of_s_g1[..len],
 * pcs/univariate/kzg.rs:28).  Least-recently-used entries are released once the cache holds
 * more than the byte limit (default: half of the device's memory). */
int plonkish_cuda_bases_cached(int device, const void *bases_affine64, size_t n, uint64_t *handle);
/* The same for plonkish_cuda_msm_bn254_g1_multi: the slice is sharded over devices 0..n_gpus-1 (exact length only). */
int plonkish_cuda_bases_cached_sharded(int n_gpus, const void *bases_affine64, size_t n, uint64_t *handle);
/* Forget (and release) the entry cached for this address: the Drop hook of the slice's owner. */
int plonkish_cuda_bases_cache_evict(const void *bases_affine64);
/* Bytes of device memory the cache may hold (0 = default); evicts down to it at once. */
int plonkish_cuda_bases_cache_limit(size_t max_bytes);
/* out[0] = cached slices, out[1] = bytes of device memory they hold. */
int plonkish_cuda_bases_cache_stats(size_t out[2]);

/* Same, from a device pointer on `device` (no host copy).  Registration expands the slice
 * into a table of window multiples T[w][i] = 2^(c*w) * P_i (W = ceil(254/c) rows, c chosen
 * from n; 12-32x the memory of the bases) so that every window shares one bucket set: no
 * per-window reduction, no 2^(c*w) doubling chain (msm.rs:162-164), and wider windows.
 * mode: 0 = policy (table when it fits in free HBM, unless PLONKISH_CUDA_PRECOMPUTE=0),
 * 1 = plain bases only (borrows d_bases, which must stay alive), 2 = table or fail.
 * plonkish_cuda_bases_register follows mode 0. */
int plonkish_cuda_bases_register_device(int device, const void *d_bases_affine64, size_t n, int mode, uint64_t *handle);

/* ---- the hot path, host buffers in, host result out -----------------------------------
 * out = sum_i scalars[i] * bases[i].  Exactly one of (bases_affine64, bases_handle != 0)
 * selects the bases; with a handle the bases are not copied again.  Runs on the
 * handle's device, or device 0.  Replaces msm.rs:84-115 for C = bn256::G1Affine.
 * Scalars (and unregistered bases) may live in pageable memory — a Rust Vec<Fr> is: such
 * uploads go through the library's pinned staging ring (copier threads + one cudaMemcpyAsync
 * per 4 MiB piece) so that they overlap the compute like uploads from pinned memory do.
 * With the timer on (plonkish_cuda_timer_config) it prints the reference's timer lines
 * "Start:/End: variable_base_msm-{n}" (msm.rs:92) in ark_std perf_trace format. */
int plonkish_cuda_msm_bn254_g1(const void *scalars_mont32, const void *bases_affine64, uint64_t bases_handle, size_t n,
                               void *out_affine64);

/* `count` MSMs of n points each against one registered base slice, results in
 * out_affine64_list[j*64 ..].  Replaces the loop of MultilinearKzg::batch_commit
 * (pcs/multilinear/kzg.rs:259-274) — same results in the same order; the upload of the
 * next polynomial's scalars overlaps the current MSM and the host blocks once. */
int plonkish_cuda_msm_bn254_g1_batch(const void *const *scalars_mont32_list, size_t count, uint64_t bases_handle, size_t n,
                                     void *out_affine64_list);

/* `count` independent MSMs, MSM j with ns[j] host scalars against the registered slice
 * bases_handles[j] (all on one device); results in out_affine64_list[j*64 ..].  The shape of
 * MultilinearKzg::open (pcs/multilinear/kzg.rs:291-293): k quotient commitments of sizes
 * 2^(k-1)..1 against eqs[k-1..0], which do not depend on one another.  Large MSMs run in turn
 * with pipelined uploads, small ones concurrently on side streams; one host wait. */
int plonkish_cuda_msm_bn254_g1_many(const void *const *scalars_mont32_list, const uint64_t *bases_handles, const size_t *ns,
                                    size_t count, void *out_affine64_list);

/* ---- callers either side of the MSM ----------------------------
 * Resident scalars: a polynomial's 2^k evaluations (n x 32 B Montgomery Fr) kept in HBM between
 * its commit and its opening.  The reference re-reads poly.evals() from host memory in every
 * caller (pcs/multilinear/kzg.rs:255, 291); here they cross PCIe once. */
int plonkish_cuda_scalars_register(int device, const void *scalars_mont32, size_t n, uint64_t *handle);
int plonkish_cuda_scalars_release(uint64_t handle);
/* Copy scalars [offset, offset + n) of a resident polynomial back to the host. */
int plonkish_cuda_scalars_read(uint64_t handle, size_t offset, size_t n, void *out_mont32);
/* Copy bases [offset, offset + n) of a registered (unsharded) slice back to the host —
 * how a device-built SRS (kzg_setup_eqs) is serialised (pcs/multilinear/kzg.rs:55-77). */
int plonkish_cuda_bases_read(uint64_t handle, size_t offset, size_t n, void *out_affine64);

/* MultilinearPolynomial::eq_xy(y) as a resident polynomial: the 2^num_vars evaluations of eq(x, y), lowest variable
 * first — what the sum-check prover state builds for a zero check (piop/sum_check/classic.rs:57-61).  Release with
 * scalars_release. */
int plonkish_cuda_eq_table(int device, const void *y_mont32, size_t num_vars, uint64_t *handle);

/* MultilinearKzg::commit on a resident polynomial (pcs/multilinear/kzg.rs:252-257):
 * variable_base_msm(first n resident scalars, first n bases of the slice) -> affine. */
int plonkish_cuda_msm_bn254_g1_resident(uint64_t scalars_handle, uint64_t bases_handle, size_t n, void *out_affine64);
/* plonkish_cuda_msm_bn254_g1_batch that also keeps every polynomial resident:
 * scalars_handles_out[j] owns a device copy of scalars_mont32_list[j] (release with
 * scalars_release).  For the witness / permutation polynomials HyperPlonk commits at
 * backend/hyperplonk.rs:201,251 and opens at :287. */
int plonkish_cuda_msm_bn254_g1_batch_keep(const void *const *scalars_mont32_list, size_t count, uint64_t bases_handle, size_t n,
                                          void *out_affine64_list, uint64_t *scalars_handles_out);
/* New resident polynomial sum_i coeffs[i] * poly_i over the first n evaluations of `count`
 * resident polynomials: the g_prime merge of pcs/multilinear.rs:203-213. */
int plonkish_cuda_fr_linear_combination(const uint64_t *scalars_handles, const void *coeffs_mont32, size_t count, size_t n,
                                        uint64_t *out_handle);
/* UnivariatePolynomial::div_rem by (X - z) on resident coefficients (poly/univariate.rs:144-168): the quotient of
 * UnivariateKzg::open (pcs/univariate/kzg.rs:281-282) and, one point after another, of the vanishing polynomials of
 * batch_open (:327).  The quotient is a new resident vector of the same length (top coefficient zero; release with
 * scalars_release), out_rem_mont32 the remainder, i.e. the polynomial's value at z. */
int plonkish_cuda_fr_div_linear(uint64_t scalars_handle, const void *z_mont32, uint64_t *out_quotient_handle, void *out_rem_mont32);
/* ---- Zeromorph<UnivariateKzg> (pcs/multilinear/zeromorph.rs), the multilinear PCS built on the univariate SRS --------
 * `quotients` (pcs/multilinear.rs:72-107) of a resident polynomial of 2^num_vars evaluations at `point`, kept in HBM:
 * *out_q_handle is a resident vector of 2^num_vars scalars holding quotient i (2^i values) at element offset 2^i (element
 * 0 is zero), out_eval_mont32 the remainder f(point).  Zeromorph::open starts from it (zeromorph.rs:149). */
int plonkish_cuda_fr_quotients(uint64_t scalars_handle, const void *point_mont32, size_t num_vars, uint64_t *out_q_handle, void *out_eval_mont32);
/* `count` independent MSMs over sub-ranges of one resident vector: MSM j = scalars [offsets[j], offsets[j] + ns[j]) against
 * the first ns[j] bases of bases_handles[j] (handles may repeat; all on the scalars' device).  With offsets[j] = ns[j] =
 * 2^j and every handle = powers_of_s_g1 it is UnivariateKzg::batch_commit_and_write over Zeromorph's quotients
 * (zeromorph.rs:150, univariate/kzg.rs:24-30) with no scalar crossing PCIe.  Results in out_affine64_list[j*64 ..]. */
int plonkish_cuda_msm_bn254_g1_many_resident(uint64_t scalars_handle, const size_t *offsets, const uint64_t *bases_handles, const size_t *ns,
                                             size_t count, void *out_affine64_list);
/* A handle on the sub-range [offset, offset + n) of a resident vector: shares the memory (no copy) and keeps it alive
 * until the slice is released (scalars_release).  One quotient or one fold as a polynomial of its own. */
int plonkish_cuda_scalars_slice(uint64_t handle, size_t offset, size_t n, uint64_t *out_handle);
/* fr_linear_combination for univariate polynomials of different lengths (`f += (scalar, q)`, poly/univariate.rs; the sums
 * of UnivariateKzg::batch_open, pcs/univariate/kzg.rs:330,339-343): a polynomial shorter than n counts as zero past its
 * last coefficient; the result has n coefficients. */
int plonkish_cuda_fr_linear_combination_padded(const uint64_t *scalars_handles, const void *coeffs_mont32, size_t count, size_t n,
                                               uint64_t *out_handle);
/* ---- Gemini<UnivariateKzg> (pcs/multilinear/gemini.rs) ------------------------------------------------------------
 * The folds of Gemini::open (gemini.rs:98-108): f_i = merge_into(f_(i-1), point[i-1], 1, 0) (poly/multilinear.rs:599-618)
 * for i = 1..num_vars-1, packed like the quotients: f_i (2^(num_vars-i) values) at element offset 2^(num_vars-i) of a new
 * resident vector of 2^num_vars scalars (elements 0 and 1 zero).  They are committed with msm_bn254_g1_many_resident
 * (gemini.rs:124-128) and opened with UnivariateKzg::batch_open through scalars_slice handles (gemini.rs:140). */
int plonkish_cuda_fr_gemini_folds(uint64_t scalars_handle, const void *point_mont32, size_t num_vars, uint64_t *out_handle);
/* q_hat of Zeromorph::open (zeromorph.rs:157-168) from the packed quotients: q_hat[2^n - 2^i + j] += weights[i] * q_i[j]
 * with weights[i] = y^i (num_vars Montgomery Fr).  A new resident vector of 2^num_vars coefficients. */
int plonkish_cuda_zeromorph_q_hat_bn254(uint64_t q_handle, const void *weights_mont32, size_t num_vars, uint64_t *out_handle);
/* f of Zeromorph::open (zeromorph.rs:175-180): f = z * poly + q_hat, f[0] += c0 (= eval_scalar * eval), f[j] +=
 * q_scalars[i] * q_i[j] for j < 2^i; poly's evaluations are read as coefficients.  A new resident vector, which
 * UnivariateKzg::open (fr_div_linear + msm_bn254_g1_resident against open_pp) opens at x (zeromorph.rs:185). */
int plonkish_cuda_zeromorph_f_bn254(uint64_t poly_handle, uint64_t q_hat_handle, uint64_t q_handle, const void *z_mont32, const void *c0_mont32,
                                    const void *q_scalars_mont32, size_t num_vars, uint64_t *out_handle);
/* permutation_z_polys (backend/hyperplonk/prover.rs:252-345), the producer of the polynomials HyperPlonk commits at
 * backend/hyperplonk.rs:251-252: per chunk of permutation polynomials the row-wise product of
 * (beta * id + gamma + value) / (beta * sigma + gamma + value), then the running product over the rows in
 * BooleanHypercube order (util/arithmetic/bh.rs:118-133), stored in index order.  value_handles[i] / sigma_handles[i]:
 * the witness column and the permutation polynomial of permutation polynomial i (count of them, resident, 2^num_vars
 * evaluations each), num_chunks = num_permutation_z_polys; out_handles receives num_chunks resident polynomials. */
int plonkish_cuda_permutation_z_polys_bn254(const uint64_t *value_handles, const uint64_t *sigma_handles, size_t count, size_t num_chunks,
                                            size_t num_vars, const void *beta_mont32, const void *gamma_mont32, uint64_t *out_handles);
/* One table of a compiled sum-check expression.  The reference's evaluator keeps identity and Lagrange polynomials,
 * constants and rotated queries implicit (piop/sum_check/classic.rs:40-75, 104-126; classic/eval.rs); a sub-expression of
 * degree <= 1 such as w + beta * (offset + identity) + gamma (backend/hyperplonk/preprocessor.rs:153-165) is a
 * multilinear polynomial itself, so it can be one explicit table with the same round values:
 *   out[b] = constant + identity_coeff * b + sum_i coeffs[i] * poly_i[rotate(b, rotations[i])],  then out[rows[j]] += values[j]
 * with rotate = BooleanHypercube::rotate (util/arithmetic/bh.rs:104-121).  poly_handles: `count` resident polynomials of
 * 2^num_vars evaluations on `device` (count may be 0: identity, Lagrange and instance polynomials are sparse_rows /
 * identity_coeff only); rotations may be null (all zero); constant / identity_coeff may be null (absent). */
int plonkish_cuda_fr_affine_table(int device, size_t num_vars, const uint64_t *poly_handles, const int32_t *rotations, const void *coeffs_mont32,
                                  size_t count, const void *constant_mont32, const void *identity_coeff_mont32, const uint64_t *sparse_rows,
                                  const void *sparse_values_mont32, size_t sparse_count, uint64_t *out_handle);
/* A compiled expression on every row: out[b] = (sum_t coeff_t * prod_j poly[fac_t,j][b]) (* poly[common][b]); the term
 * encoding is plonkish_cuda_sumcheck_new's.  The compressed input / table polynomial of a lookup argument
 * (lookup_compressed_poly, backend/hyperplonk/prover.rs:79-137) is one call. */
int plonkish_cuda_fr_expression_table(const uint64_t *poly_handles, size_t num_polys, size_t num_vars, const void *term_coeffs_mont32,
                                      const uint32_t *term_offsets, const uint32_t *term_polys, size_t num_terms, int common_poly,
                                      uint64_t *out_handle);
/* lookup_m_poly (backend/hyperplonk/prover.rs:145-192): m[row] = number of input values equal to table[row]; a value
 * present in several table rows is counted at the last of them (the reference's HashMap, :151).  Returns
 * PLONKISH_CUDA_E_INVALID ("Invalid lookup input", :176-178) when an input value is not in the table. */
int plonkish_cuda_lookup_m_poly_bn254(uint64_t input_handle, uint64_t table_handle, uint64_t *out_handle);
/* lookup_h_poly (backend/hyperplonk/prover.rs:206-250): h = 1 / (gamma + input) - m / (gamma + table), resident in and out. */
int plonkish_cuda_lookup_h_poly_bn254(uint64_t input_handle, uint64_t table_handle, uint64_t m_handle, const void *gamma_mont32,
                                      uint64_t *out_handle);
/* MultilinearPolynomial::evaluate (poly/multilinear.rs:137-156) of a resident polynomial at `count` points of num_vars
 * Montgomery Fr each (points_mont32: count x num_vars x 32 B): the evaluations prove_sum_check needs at the rotated
 * points (backend/hyperplonk/prover.rs:392-400 through evaluate_for_rotation, poly/multilinear.rs:191-263).
 * out_evals_mont32: count x 32 B. */
int plonkish_cuda_fr_evaluate(uint64_t scalars_handle, const void *points_mont32, size_t num_vars, size_t count, void *out_evals_mont32);
/* MultilinearKzg::open on a resident polynomial of 2^num_vars evaluations
 * (pcs/multilinear/kzg.rs:276-302): `quotients` (pcs/multilinear.rs:72-107) runs in HBM and
 * the num_vars quotient MSMs (kzg.rs:291-293) read their scalars from there.  eq_handles[i] =
 * the registered slice pp.eq(i) (2^i bases), i < num_vars.  out_comms_affine64 receives
 * num_vars x 64 B in the order kzg.rs:299 writes them to the transcript; out_eval_mont32
 * receives f(point), the `remainder` of kzg.rs:295. */
int plonkish_cuda_kzg_open_bn254(uint64_t scalars_handle, const uint64_t *eq_handles, const void *point_mont32, size_t num_vars,
                                 void *out_comms_affine64, void *out_eval_mont32);

/* fixed_base_msm (util/arithmetic/msm.rs:67-81) over the window table of one base (msm.rs:16-31)
 * followed by batch_normalize (pcs/multilinear/kzg.rs:204-207, pcs/univariate/kzg.rs:196-199):
 * out_affine64_list[i] = scalars[i] * base.  Host in, host out; the table lives on the device
 * for the duration of the call (signed 16-bit windows, 32 MiB). */
int plonkish_cuda_fixed_base_msm_bn254_g1(int device, const void *base_affine64, const void *scalars_mont32, size_t n,
                                          void *out_affine64_list);
/* The prover half of MultilinearKzg::setup (pcs/multilinear/kzg.rs:167-212) on the device:
 * eq tables from ss (num_vars Montgomery Fr, kzg.rs:174-193), times g1 by fixed-base MSM and
 * normalised (:195-208), every slice eqs[k] (2^k bases, k = 0..num_vars) registered as resident
 * bases.  handles_out receives num_vars + 1 handles (release each with bases_release; read
 * back with bases_read).  The SRS never exists in host memory. */
int plonkish_cuda_kzg_setup_eqs_bn254(int device, const void *g1_affine64, const void *ss_mont32, size_t num_vars,
                                      uint64_t *handles_out);

/* The G1 half of UnivariateKzg::setup (pcs/univariate/kzg.rs:175-195) on the device: powers(s).take(n), times g1 by
 * fixed-base MSM, normalised, registered as one resident slice (powers_of_s_g1, the bases of commit_coeffs,
 * pcs/univariate/kzg.rs:24-30; any prefix can be used, as trim does at :217-229).  Read back with bases_read. */
int plonkish_cuda_kzg_setup_powers_bn254(int device, const void *g1_affine64, const void *s_mont32, size_t n, uint64_t *handle);

/* ---- sum check ---------------------------------------------------------
 * ClassicSumCheck<EvaluationsProver>::prove (piop/sum_check/classic.rs:208-240) round by round.  The
 * caller owns the transcript: it writes each round message, squeezes the challenge (classic.rs:226-229)
 * and hands it back; the tables stay in HBM.
 *
 * The expression is flattened: sum_t coeffs[t] * prod_{j in [offsets[t], offsets[t+1])} poly[term_polys[j]],
 * times poly[common_poly] when common_poly >= 0 (the eq(x, y) of a zero check).  Every polynomial is a
 * resident table of 2^num_vars evaluations (scalars_register / batch_keep / fr_linear_combination); eq_xy,
 * identity and Lagrange tables and rotated copies, which the reference keeps implicit (classic.rs:40-75,
 * 104-126), are materialised by the caller.  Limits: 32 polynomials, 32 terms, 8 factors per term, degree 8.
 * The resident polynomials are not modified. */
int plonkish_cuda_sumcheck_new(const uint64_t *poly_handles, size_t num_polys, size_t num_vars, const void *term_coeffs_mont32,
                               const uint32_t *term_offsets, const uint32_t *term_polys, size_t num_terms, int common_poly,
                               uint64_t *state_handle);
/* Degree of the round polynomial: max factors per term (+ 1 with a common factor); negative on error. */
int plonkish_cuda_sumcheck_degree(uint64_t state_handle);
/* The round message without its first entry (piop/sum_check/classic/eval.rs:101-131):
 * out_evals_mont32[x-1] = sum_b expr(r_0, .., r_{round-1}, X = x, b) for x = 1..degree, over the pairs
 * (2b, 2b+1) of every table (eval.rs:236-243).  The caller sets evals[0] = sum - evals[1] (eval.rs:128). */
int plonkish_cuda_sumcheck_round(uint64_t state_handle, void *out_evals_mont32);
/* The same round for a zero check whose common factor is eq(x, y) (times a constant), factored: the round polynomial is
 * h(X) = (1 - y_r + X (2 y_r - 1)) * G(X) with G of one degree less; out_evals receives G(1..degree-1), computed with the
 * common factor's pair sum instead of its walk — one evaluation point fewer per pair.  The caller rebuilds the message
 * h(0..degree) of eval.rs:101-131 from G and the running sum (G(0) from h(0) + h(1) = sum); plonkish_b200/sumcheck.py. */
int plonkish_cuda_sumcheck_round_factored(uint64_t state_handle, void *out_evals_mont32);
/* ProverState::next_round (classic.rs:90-141): fix the lowest variable of every table at the challenge
 * (MultilinearPolynomial::fix_var, poly/multilinear.rs:179-189: out[b] = (e[2b+1] - e[2b]) * x + e[2b]). */
int plonkish_cuda_sumcheck_fix_var(uint64_t state_handle, const void *challenge_mont32);
/* ProverState::into_evals (classic.rs:143-149) after num_vars rounds: num_polys field elements. */
int plonkish_cuda_sumcheck_final_evals(uint64_t state_handle, void *out_evals_mont32);
int plonkish_cuda_sumcheck_free(uint64_t state_handle);

/* Keccak-f[1600] on 25 little-endian lanes (lane (x, y) at index x + 5 y), host code: the permutation of the reference's
 * Keccak256Transcript (util/transcript.rs:100-131; util/hash.rs:5-8 takes Keccak256 from the sha3 crate).  The host
 * mirrors build the sponge and the transcript rules on top of it. */
void plonkish_cuda_keccak_f1600(uint64_t state[25]);

/* Same as plonkish_cuda_msm_bn254_g1 for the reference's non-contiguous callers, which pass iterators of
 * references (chain![..] at pcs/univariate/kzg.rs:346,408; .map(|c| &c.0) at
 * pcs/multilinear/kzg.rs:145): gathers the n scalars and n bases into staging first. */
int plonkish_cuda_msm_bn254_g1_gather(const void *const *scalar_ptrs, const void *const *base_ptrs, size_t n,
                                      void *out_affine64);

/* Point-sharded over the first n_gpus devices of this process: GPU g runs the
 * whole pipeline on points [g*ceil(n/G), (g+1)*ceil(n/G)) — one host thread per device, the
 * shard's scalars uploaded in chunks overlapped with its compute, as in the single-GPU call —
 * the G 128-byte projective partials are gathered with ncclAllGather (ncclCommInitAll
 * communicator, NVLink) and device 0 adds and normalises them.  This is msm.rs:101-114 (chunk
 * per thread, fold the partials) lifted from threads to GPUs.  bases_affine64 is uploaded per
 * call unless bases_handle names a slice registered with .._register_sharded.  NCCL is loaded at
 * run time (libnccl.so.2); without it the call fails with PLONKISH_CUDA_E_COMM. */
int plonkish_cuda_bases_register_sharded(int n_gpus, const void *bases_affine64, size_t n, uint64_t *handle);
/* The same from device memory: d_bases_affine64[g] points at shard g (ceil(n/G) points; trailing shards may be short
 * or empty) on device g; mode as in plonkish_cuda_bases_register_device. */
int plonkish_cuda_bases_register_sharded_device(int n_gpus, const void *const *d_bases_affine64, size_t n, int mode, uint64_t *handle);
int plonkish_cuda_msm_bn254_g1_multi(int n_gpus, const void *scalars_mont32, const void *bases_affine64,
                                     uint64_t bases_handle, size_t n, void *out_affine64);

/* ---- device-resident entry points (no host copies, stream-ordered) ---------------------
 * d_scalars / d_bases are device pointers on `device`; the call only enqueues work on
 * `cuda_stream` (a cudaStream_t; NULL = the legacy default stream, as everywhere in the
 * CUDA runtime) and returns.
 * window_bits = 0 picks the default for n, 8..16 forces it.  Either output may be NULL:
 * d_out_affine64 receives the normalised point, d_out_xyzz128 the projective partial
 * (what one rank contributes before the multi-GPU gather). */
int plonkish_cuda_msm_bn254_g1_device(int device, const void *d_scalars, const void *d_bases, size_t n,
                                      uint32_t window_bits, void *d_out_affine64, void *d_out_xyzz128, void *cuda_stream);

/* Device-resident scalars against a registered (single-device) base slice; runs on the
 * handle's device with the table of window multiples when the handle has one. */
int plonkish_cuda_msm_bn254_g1_device_resident(const void *d_scalars, uint64_t bases_handle, size_t n, void *d_out_affine64,
                                               void *d_out_xyzz128, void *cuda_stream);

/* Host scalars against a registered slice, projective partial left in device memory
 * (d_out_xyzz128): what one rank of the one-process-per-GPU deployment computes before the
 * all-gather.  Blocking; the scalar upload is pipelined like plonkish_cuda_msm_bn254_g1. */
int plonkish_cuda_msm_bn254_g1_host_partial(const void *scalars_mont32, uint64_t bases_handle, size_t n, void *d_out_xyzz128);

/* Adds `count` projective partials (128 B each, e.g. one per rank after an NCCL
 * all-gather) and normalises: msm.rs:112-114 followed by the caller's to_affine(). */
int plonkish_cuda_g1_sum_partials_device(int device, const void *d_partials_xyzz128, size_t count, void *d_out_affine64,
                                         void *cuda_stream);

/* ---- introspection / measurement helpers ------------------------------------------------- */

/* Fills out[0..8) with the plan the library would use for n points on `device`:
 * window bits c, windows W, high/low bucket bits of the two sort levels, point
 * index bits, points per decompose tile, run length per accumulate thread, and the
 * number of accumulate threads.  With a bases_handle the plan is the handle's (table) plan. */
int plonkish_cuda_msm_plan(int device, size_t n, uint32_t window_bits, uint64_t bases_handle, uint32_t out[8]);

/* One synchronous MSM on the context stream with CUDA events between the stages;
 * stage_ms[0..9) = decompose, scans, bin scatter, bin sort, accumulate, item levels,
 * bucket reduce, window combine, finalize (to_affine).  d_out_affine64 may be NULL; bases come
 * from d_bases or, when bases_handle != 0, from the registered slice. */
int plonkish_cuda_msm_profile_device(int device, const void *d_scalars, const void *d_bases, uint64_t bases_handle, size_t n,
                                     uint32_t window_bits, void *d_out_affine64, double stage_ms[9]);

/* Kernels launched by the library in this process since init (for bench accounting). */
uint64_t plonkish_cuda_launch_count(void);
/* Bytes uploaded through the pinned staging ring so far (pageable sources). */
uint64_t plonkish_cuda_staged_bytes(void);
/* GB/s the staging ring sustained over its recent large uploads (0 before the first one); the host-scalar MSM cuts its
 * points into more chunks when this falls below what its three-chunk pipeline needs. */
double plonkish_cuda_staging_rate_gbps(void);

/* The reference's timer lines (msm.rs:92 -> util/timer.rs:19-24 -> ark_std perf_trace; parsed by
 * benchmark/src/bin/plotter.rs:337-373).  mode: 0 off, 1 stderr, 2 stdout (also PLONKISH_CUDA_TIMER=1 /
 * =stdout); depth: nesting depth of the caller's open timers (indentation of the lines; also
 * PLONKISH_CUDA_TIMER_DEPTH).  Every entry point prints one Start/End pair per MSM it performs, named
 * variable_base_msm-{n}: a single call with its wall time; the batch entry with each MSM's own span on the
 * compute stream; the many / open entries, whose MSMs run concurrently, with the call's wall time split by
 * the library's load model — in every case the lines of one call add up to the call's duration, which is
 * what the plotter's cost breakdown needs (it subtracts them from the enclosing timer). */
int plonkish_cuda_timer_config(int mode, int depth);
/* One Start/End pair "variable_base_msm-{n}" with the given duration, in the configured mode (for a host mirror that
 * composes an MSM out of several calls and times it itself). */
void plonkish_cuda_timer_emit(size_t n, double ms);

/* FP64-pipe probe: out[0] = independent fma.rz.f64 per second; out[1] = DFMA per second and out[2] =
 * mad.wide.u32 per second when the two are interleaved 1:1 in one instruction stream. */
int plonkish_cuda_bench_fp64_pipe(int device, double out[3]);
/* Issue probe: out[R] = independent mad.wide.u32 per second when R independent 32-bit adds (ALU pipe) are issued next to
 * every multiply, R = 0..3: flat = the multiplier pipe is the bound, falling = instruction issue is. */
int plonkish_cuda_bench_issue_mix(int device, double out[4]);
/* Register-resident mixed-addition streams: out[0] = additions per second of the FP64-pipe formulas (dpfq.cuh) alone with
 * dp_blocks_per_sm blocks of 128 threads per SM, out[1] = of the integer-pipe formulas alone with int_blocks_per_sm blocks
 * per SM, out[2] / out[3] = of the same two kernels running at the same time on two streams, out[4] = wall ms of that run. */
int plonkish_cuda_bench_dp_madd(int device, int dp_blocks_per_sm, int int_blocks_per_sm, double out[5]);

/* Integer-pipe microbenchmarks on `device` (CUDA-event timed):
 *   out[0] = independent mad.wide.u32 (IMAD.WIDE.U32) per second, all SMs busy
 *   out[1] = 254-bit Montgomery products per second from the library's own fq_mul
 *   out[2] = the device's maximum SM clock in MHz, out[3] = SM count
 *   out[4] = independent 32-bit mad.lo.u32 (IMAD) per second
 *   out[5] = wide multiply-adds per second inside mad.lo.cc/madc.hi.cc carry chains
 *            (IMAD.WIDE.U32.X, the instruction the Montgomery products are built from) */
int plonkish_cuda_bench_integer_pipe(int device, double out[6]);
/* Field products per second inside register-resident mixed-addition streams shaped like
 * k_accumulate (128-thread blocks, no memory traffic): out[0] one XYZZ accumulator per thread at
 * 128 registers (4 blocks/SM), out[1] two accumulators per thread, out[2] one dependent fq_mul
 * chain per thread, out[3] / out[4] = out[0] / out[1] with the register cap lifted (2 blocks/SM).
 * k_accumulate's ceiling is out[0]; the fq_mul-stream peak of bench_integer_pipe is what two
 * short independent chains per thread reach at 30 registers. */
int plonkish_cuda_bench_madd(int device, double out[5]);
/* Rows of 8 partial products (8 limbs x 1 limb, accumulated) per second: out[0] spelled as 16
 * 32-bit IMADs in two carry chains (lo then hi), out[1] as 8 fused IMAD.WIDE (the library's form). */
int plonkish_cuda_bench_row_forms(int device, double out[2]);
/* Field inversions per second: out[0] safegcd division steps (used), out[1] Fermat ladder. */
int plonkish_cuda_bench_inversion(int device, double out[2]);
/* The library's fq_mul stream with warps_per_sm (multiple of 4, 4..64) resident warps per SM. */
int plonkish_cuda_bench_fq_mul_occupancy(int device, int warps_per_sm, double *out_per_s);

/* Synthetic bases with a known discrete log: d_out[i] = (a + i*step) * G for
 * i in [first, first + n), affine, written on the device.  Used by bench.py and the
 * tests so that sum_i s_i*B_i can be checked against ((a*sum s_i + step*sum i*s_i) mod r)*G
 * at any size.  The reference's SRS has the same shape
 * (eq_j(s) * G, pcs/multilinear/kzg.rs:174-212). */
int plonkish_cuda_synth_bases_device(int device, void *d_out_affine64, size_t first, size_t n, uint64_t a, uint64_t step,
                                     void *cuda_stream);

/* ---- test hooks ------------------------------------------------------------------------------
 * Element-wise probes of the device arithmetic on host arrays, for the parity tests.
 * field ops (32-byte elements): 0 Fq mul, 1 Fq add, 2 Fq sub, 3 Fr Montgomery->canonical
 * (halo2_curves to_repr, msm.rs:153), 4 Fq inverse (Fermat ladder), 5 Fr mul, 6 Fq negate,
 * 7 Fq inverse (safegcd, the one the library uses), 8 Fq fused a[i]*b[i] + b[i]*a[(i+1)%n] with one
 * reduction (the y-coordinate form of the point formulas), 9 the same shape in Fr: a[i]^2 + b[i]*b[(i+1)%n],
 * 10 Fq square (symmetric partial products taken once), 11 Fr square, 12 Fq product and 13 Fq square on the FP64 pipe
 * (dpfq.cuh: six 48-bit limbs in doubles, round-toward-zero DFMAs), 14 words -> double limbs -> words.
 * point ops (128-byte X,Y,ZZ,ZZZ slots; b's first 64 bytes are an affine point for op 0):
 * 0 mixed add, 1 full add, 2 double, 3 to_affine (result in the first 64 bytes), 4 mixed add through the FP64-pipe formulas. */
int plonkish_cuda_debug_field_op(int device, int op, const void *a32, const void *b32, void *out32, size_t n);
int plonkish_cuda_debug_point_op(int device, int op, const void *a128, const void *b128, void *out128, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* PLONKISH_CUDA_H */
