// tests/emul/emul_msm.cpp — TEST INFRASTRUCTURE ONLY.
// Builds the product's kernel sources for the CPU through cuda_emul.h and exposes
// the launch sequence plus a few stage probes to pytest (tests/test_emulated_kernels.py).
#define PLONKISH_EMUL 1
#include "cuda_emul.h"

#include "../../plonkish_b200/csrc/msm_kernels.cuh"
#include "../../plonkish_b200/csrc/poly_kernels.cuh"
#include "../../plonkish_b200/csrc/sumcheck_kernels.cuh"
#include "../../plonkish_b200/csrc/lookup_kernels.cuh"
#include "../../plonkish_b200/csrc/dpfq.cuh"

#include <stdlib.h>

#include <vector>

using namespace pk;

extern "C" {

// FP64-pipe arithmetic (dpfq.cuh): fma() under FE_TOWARDZERO stands for __fma_rz.
// op 0: out = a * b / 2^288 (raw limbs), 1: a^2 / 2^288, 2..4: a - b + {2, 4, 8} p, 5: a + b.  Limbs are 6 x u64 < 2^48.
void emul_dp_raw(int op, const uint64_t *a6, const uint64_t *b6, uint64_t *o6) {
    DpRoundGuard guard;
    dfe a, b, r;
    for (int i = 0; i < 6; ++i) { a.l[i] = (double)a6[i]; b.l[i] = (double)b6[i]; }
    switch (op) {
        case 0: r = dp_mul(a, b); break;
        case 1: r = dp_sqr(a); break;
        case 2: r = dp_sub_fe<2>(a, b); break;
        case 3: r = dp_sub_fe<4>(a, b); break;
        case 4: r = dp_sub_fe<8>(a, b); break;
        default: r = dp_add_fe(a, b); break;
    }
    for (int i = 0; i < 6; ++i) o6[i] = (r.l[i] >= 0 && r.l[i] == (double)(uint64_t)r.l[i]) ? (uint64_t)r.l[i] : ~0ull;
}
// memory form in, memory form out: must equal fq_mul bit for bit.
void emul_dp_mul_words(const u32 *a, const u32 *b, u32 *o) {
    DpRoundGuard guard;
    fe x, y;
    memcpy(&x, a, 32); memcpy(&y, b, 32);
    fe r = dp_to_mont256(dp_mul(dp_from_mont256(x), dp_from_mont256(y)));
    memcpy(o, &r, 32);
}
void emul_dp_words_roundtrip(const u32 *a, u32 *o) {
    DpRoundGuard guard;
    fe x;
    memcpy(&x, a, 32);
    fe r = dp_to_words(dp_from_words(x));
    memcpy(o, &r, 32);
}
// count mixed additions of pts (affine, memory form) into acc through the FP64-pipe formulas; acc in / out as XYZZ words.
void emul_dxyzz_madd(u32 *acc128, const u32 *pts64, u32 count) {
    DpRoundGuard guard;
    xyzz a;
    memcpy(&a, acc128, 128);
    dxyzz d = dxyzz_from_words(a);
    for (u32 k = 0; k < count; ++k) {
        affine p;
        memcpy(&p, pts64 + 16 * k, 64);
        dxyzz_madd(d, p.x, p.y);
    }
    a = dxyzz_to_words(d);
    memcpy(acc128, &a, 128);
}

// permutation_z_polys: values / sigmas are `count` arrays of 2^k elements; out receives num_chunks arrays of 2^k elements.
void emul_permutation_z(const void *const *values, const void *const *sigmas, u32 count, u32 num_chunks, u32 k, const void *beta, const void *gamma, void *out) {
    const size_t n = (size_t)1 << k;
    const size_t chunk_size = (count + num_chunks - 1) / num_chunks;
    std::vector<unsigned char> scratch(pk_perm_z_scratch_elems(num_chunks, n) * 32), bg(64);
    memcpy(bg.data(), beta, 32);
    memcpy(bg.data() + 32, gamma, 32);
    std::vector<PermArgs> chunks(num_chunks);
    std::vector<void *> outs(num_chunks);
    for (u32 c = 0; c < num_chunks; ++c) {
        PermArgs &a = chunks[c];
        memset(&a, 0, sizeof(a));
        const size_t first = c * chunk_size;
        a.count = (u32)(count - first < chunk_size ? count - first : chunk_size);
        for (u32 i = 0; i < a.count; ++i) {
            a.value[i] = (const uint4 *)values[first + i];
            a.sigma[i] = (const uint4 *)sigmas[first + i];
            a.id_offset[i] = (unsigned long long)(first + i) << k;
        }
        outs[c] = (char *)out + (size_t)c * n * 32;
    }
    pk_enqueue_perm_z(chunks.data(), num_chunks, k, bg.data(), scratch.data(), outs.data(), 0);
}

// div_rem by (X - z): q receives n elements (q[n-1] = 0), rem one.
void emul_div_linear(const u32 *c, uint64_t n, const u32 *z, u32 *q, u32 *rem) {
    std::vector<unsigned char> scratch((pk_horner_scratch_elems(n) + 16) * 32);
    memcpy(scratch.data(), z, 32);
    pk_enqueue_div_linear(c, n, scratch.data(), scratch.data() + 16 * 32, q, rem, 0);
}
// the same with 2^log_chunk coefficients per thread (the library's PLONKISH_CUDA_HORNER_LOG_CHUNK)
void emul_div_linear_chunk(const u32 *c, uint64_t n, const u32 *z, u32 *q, u32 *rem, u32 log_chunk) {
    std::vector<unsigned char> scratch((pk_horner_scratch_elems(n, log_chunk) + 16) * 32);
    memcpy(scratch.data(), z, 32);
    pk_enqueue_div_linear(c, n, scratch.data(), scratch.data() + 16 * 32, q, rem, 0, log_chunk);
}

// redundant-range forms of the accumulate loop: op 0 mul, 1 sqr, 2 add, 3 sub, 4 neg, 5 mul_sum (a*b + c*d); inputs in [0, 2p)
void emul_fq_lazy(int op, const u32 *a, const u32 *b, const u32 *c, const u32 *d, u32 *o) {
    fe x, y, z, w, r;
    memcpy(&x, a, 32); memcpy(&y, b, 32); memcpy(&z, c, 32); memcpy(&w, d, 32);
    switch (op) {
        case 0: r = fq_mul_lz(x, y); break;
        case 1: r = fq_sqr_lz(x); break;
        case 2: r = fq_add_lz(x, y); break;
        case 3: r = fq_sub_lz(x, y); break;
        case 4: r = fq_neg_lz(x); break;
        default: r = fq_mul_sum_lz(x, y, z, w); break;
    }
    memcpy(o, &r, 32);
}
void emul_xyzz_madd_lazy(u32 *acc128, const u32 *pts64, u32 count) {
    xyzz a;
    memcpy(&a, acc128, 128);
    for (u32 k = 0; k < count; ++k) {
        affine p;
        memcpy(&p, pts64 + 16 * k, 64);
        xyzz_madd_lazy(a, p.x, p.y);
    }
    a = xyzz_canonical(a);
    memcpy(acc128, &a, 128);
}

// Field / point probes (the portable restatements of the carry-chain blocks).
void emul_fq_mul(const u32 *a, const u32 *b, u32 *o) {
    fe x, y;
    memcpy(&x, a, 32); memcpy(&y, b, 32);
    fe r = fq_mul(x, y);
    memcpy(o, &r, 32);
}
void emul_fq_sqr(const u32 *a, u32 *o) {
    fe x;
    memcpy(&x, a, 32);
    fe r = fq_sqr(x);
    memcpy(o, &r, 32);
}
void emul_fq_mul_sum(const u32 *a, const u32 *b, const u32 *c, const u32 *d, u32 *o) {
    fe x, y, z, w;
    memcpy(&x, a, 32); memcpy(&y, b, 32); memcpy(&z, c, 32); memcpy(&w, d, 32);
    fe r = fq_mul_sum(x, y, z, w);
    memcpy(o, &r, 32);
}
void emul_fr_to_canonical(const u32 *a, u32 *o) {
    fe x;
    memcpy(&x, a, 32);
    fe r = fr_to_canonical(x);
    memcpy(o, &r, 32);
}
void emul_xyzz_madd(u32 *acc128, const u32 *pt64) {
    xyzz a; affine p;
    memcpy(&a, acc128, 128); memcpy(&p, pt64, 64);
    xyzz_madd(a, p.x, p.y);
    memcpy(acc128, &a, 128);
}
void emul_xyzz_add(const u32 *a128, const u32 *b128, u32 *o128) {
    xyzz a, b;
    memcpy(&a, a128, 128); memcpy(&b, b128, 128);
    xyzz r = xyzz_add(a, b);
    memcpy(o128, &r, 128);
}
void emul_xyzz_to_affine(const u32 *a128, u32 *o64) {
    xyzz a;
    memcpy(&a, a128, 128);
    affine r = xyzz_to_affine(a);
    memcpy(o64, &r, 64);
}

void emul_plan(u32 n, u32 c, u32 sms, u32 *out /*[8]*/) {
    MsmPlan p = pk_make_plan(n, c, sms);
    out[0] = p.c; out[1] = p.W; out[2] = p.hi_bits; out[3] = p.lo_bits; out[4] = p.idx_bits;
    out[5] = p.tile; out[6] = p.L; out[7] = p.nthreads1;
}

// K3's run of every thread for `total` entries: s_out/e_out[nthreads] (s == e == 0xffffffff: no run).  Returns the
// thread count pk_acc_threads sizes for `entries` (the largest total the launch is planned for).
u32 emul_acc_runs(unsigned long long entries, u32 total, u32 tiers, u32 resident, u32 *s_out, u32 *e_out, u32 cap) {
    const u32 nthreads = pk_acc_threads(entries, tiers, resident);
    for (u32 t = 0; t < nthreads && t < cap; ++t) {
        u32 s = 0, e = 0;
        if (pk_acc_run(t, total, tiers, resident, nthreads, 0, s, e)) { s_out[t] = s; e_out[t] = e; } else { s_out[t] = e_out[t] = 0xffffffffu; }
    }
    return nthreads;
}

// Full launch sequence on host memory.  digits_out (optional): [W][n_pad] u16;
// sorted_out / bucket_start_out (optional) sized n*W and nbuckets+1.
int emul_msm(const void *scalars, const void *bases, u32 n, u32 c_override, u32 sm_count, u32 serial_items, void *out_affine64,
             u16 *digits_out, u32 *sorted_out, u32 *bucket_start_out) {
    if (n == 0) { memset(out_affine64, 0, 64); return 0; }
    MsmPlan p = pk_make_plan(n, c_override, sm_count);
    if (serial_items) p.serial_items = serial_items;
    p.blk = 32;  // fewer OS threads per emulated block; the kernels index by blockDim.x
    size_t bytes = pk_workspace_bytes(p);
    void *arena = aligned_alloc(256, bytes);
    memset(arena, 0xA5, bytes);  // poison: nothing may rely on zeroed scratch
    MsmWorkspace ws = pk_carve_workspace(p, arena);
    pk_enqueue_msm(p, scalars, bases, ws, nullptr, 0);
    affine out;
    PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, 0, ws.result, 1u, &out, (xyzz *)nullptr);
    memcpy(out_affine64, &out, 64);
    if (digits_out) memcpy(digits_out, ws.digits, sizeof(u16) * (size_t)p.W * p.n_pad);
    if (sorted_out) memcpy(sorted_out, ws.sorted, sizeof(u32) * (size_t)p.n * p.W);
    if (bucket_start_out) memcpy(bucket_start_out, ws.bucket_start, sizeof(u32) * (p.nbuckets + 1));
    free(arena);
    return 0;
}

// Mode 1: expand the first n_table bases into the table of window multiples with the
// product's table kernels, then run an MSM over the first n_use points through it.
int emul_msm_table_chunked(const void *scalars, const void *bases, u32 n_table, u32 n_use, u32 c, u32 sm_count, u32 nchunks,
                           void *out_affine64);

int emul_msm_table(const void *scalars, const void *bases, u32 n_table, u32 n_use, u32 c, u32 sm_count, void *out_affine64) {
    return emul_msm_table_chunked(scalars, bases, n_table, n_use, c, sm_count, 1, out_affine64);
}

// The host path's pipelining: the points are cut into nchunks consecutive chunks that all
// add into one bucket array (p.accum), and one reduce phase follows.
int emul_msm_table_chunked(const void *scalars, const void *bases, u32 n_table, u32 n_use, u32 c, u32 sm_count, u32 nchunks,
                           void *out_affine64) {
    if (n_use == 0) { memset(out_affine64, 0, 64); return 0; }
    if (!c) c = pk_table_window_bits(n_table);
    const u32 W = pk_windows_for(c);
    affine *table = (affine *)aligned_alloc(256, sizeof(affine) * (size_t)W * n_table);
    xyzz *cur = (xyzz *)aligned_alloc(256, sizeof(xyzz) * (size_t)n_table);
    pk_enqueue_table_build(bases, n_table, c, W, cur, table, 0);
    const u32 per = (n_use + nchunks - 1) / nchunks;
    MsmPlan p0 = pk_make_plan_b(per, c, n_table, sm_count);
    p0.blk = 32;
    p0.blk_stage = 32;
    p0.nchunks = nchunks;
    size_t bytes = pk_workspace_bytes(p0);
    void *arena = aligned_alloc(256, bytes);
    memset(arena, 0xA5, bytes);
    MsmWorkspace ws = pk_carve_workspace(p0, arena);
    MsmPlan last = p0;
    for (u32 done = 0, k = 0; done < n_use; done += per, ++k) {
        const u32 cnt = (n_use - done < per) ? n_use - done : per;
        MsmPlan p = pk_make_plan_b(cnt, c, n_table, sm_count);
        p.blk = 32;
        p.blk_stage = 32;
        p.chunk = k;
        p.nchunks = nchunks;
        MsmWorkspace w = pk_carve_workspace(p, arena);
        pk_enqueue_buckets(p, (const char *)scalars + (size_t)done * 32, (const char *)table + (size_t)done * 64, w, 0);
        last = p;
    }
    pk_enqueue_reduce(last, ws, nullptr, 0);
    affine out;
    PK_LAUNCH(k_finalize, dim3(1), dim3(32), 0, 0, ws.result, 1u, &out, (xyzz *)nullptr);
    memcpy(out_affine64, &out, 64);
    free(arena); free(cur); free(table);
    return 0;
}

// ---- the callers either side of the MSM (poly_kernels.cuh)
// quotients: q_out has 2^k scalars (q_i at offset 2^i), eval_out one.
void emul_quotients(const void *poly, u32 k, const void *point, void *q_out, void *eval_out) {
    const size_t n = (size_t)1 << k;
    std::vector<fe> ra(n / 2 + 1), rb(n / 4 + 1);
    pk_enqueue_quotients(poly, k, point, q_out, ra.data(), rb.data(), eval_out, 4, 0);
}
// Zeromorph's q_hat and f from the packed quotients (weights: num_vars x 32 B)
void emul_zm_q_hat(const void *q, const void *weights, u32 num_vars, void *out) {
    ZmWeights w;
    memset(&w, 0, sizeof(w));
    memcpy(w.w, weights, (size_t)num_vars * 32);
    pk_enqueue_zm_q_hat(q, w, num_vars, out, 1, 0);
}
void emul_zm_f(const void *poly, const void *q_hat, const void *q, const void *weights, const void *z, const void *c0, u32 num_vars, void *out) {
    ZmWeights w;
    memset(&w, 0, sizeof(w));
    memcpy(w.w, weights, (size_t)num_vars * 32);
    fe zz, cc;
    memcpy(zz.l, z, 32);
    memcpy(cc.l, c0, 32);
    pk_enqueue_zm_f(poly, q_hat, q, w, zz, cc, num_vars, out, 1, 0);
}
void emul_fr_lincomb(const void *const *polys, const void *coeffs, u32 count, u32 n, void *out) {
    for (u32 done = 0; done < count; done += PK_LINCOMB_MAX) {
        LincombArgs a;
        a.count = count - done < PK_LINCOMB_MAX ? count - done : PK_LINCOMB_MAX;
        a.accumulate = done ? 1u : 0u;
        for (u32 i = 0; i < a.count; ++i) {
            a.poly[i] = (const uint4 *)polys[done + i];
            a.len[i] = n;
            memcpy(a.coeff[i].l, (const char *)coeffs + (size_t)(done + i) * 32, 32);
        }
        PK_LAUNCH(k_fr_lincomb, dim3(3), dim3(64), 0, 0, a, (size_t)n, (uint4 *)out);
    }
}
// polynomials of lens[i] values each, zero past them; out has n values
void emul_fr_lincomb_padded(const void *const *polys, const uint64_t *lens, const void *coeffs, u32 count, u32 n, void *out) {
    for (u32 done = 0; done < count; done += PK_LINCOMB_MAX) {
        LincombArgs a;
        a.count = count - done < PK_LINCOMB_MAX ? count - done : PK_LINCOMB_MAX;
        a.accumulate = done ? 1u : 0u;
        for (u32 i = 0; i < a.count; ++i) {
            a.poly[i] = (const uint4 *)polys[done + i];
            a.len[i] = lens[done + i];
            memcpy(a.coeff[i].l, (const char *)coeffs + (size_t)(done + i) * 32, 32);
        }
        PK_LAUNCH(k_fr_lincomb, dim3(3), dim3(64), 0, 0, a, (size_t)n, (uint4 *)out);
    }
}
// Gemini's folds, packed: out has 2^num_vars scalars, f_i at element offset 2^(num_vars - i)
void emul_gemini_folds(const void *poly, u32 num_vars, const void *point, void *out) { pk_enqueue_gemini_folds(poly, num_vars, point, out, 1, 0); }
void emul_fr_powers(const void *s, u32 n, void *out) { pk_enqueue_fr_powers(s, n, out, 0); }
void emul_eq_scalars(const void *ss, u32 num_vars, void *out) { pk_enqueue_eq_scalars(ss, num_vars, out, 2, 0); }
// out[i] = scalars[i] * base through the signed-window table.
void emul_fixed_base(const void *base64, const void *scalars, u32 n, void *out_affine) {
    const size_t entries = (size_t)PK_FIXED_W * PK_FIXED_ROW;
    std::vector<affine> table(entries), offsets(PK_FIXED_W);
    std::vector<xyzz> tmp(entries > n ? entries : n);
    pk_enqueue_fixed_table(base64, offsets.data(), tmp.data(), table.data(), 0);
    pk_enqueue_fixed_base(scalars, n, table.data(), tmp.data(), (affine *)out_affine, 0);
}

// ---- sum-check rounds (sumcheck_kernels.cuh): tables of n evaluations each, flattened expression
void emul_sumcheck_round(const void *const *polys, u32 num_polys, u32 n, const void *coeffs, const u32 *offsets, const u32 *term_polys,
                         u32 num_terms, int common, u32 degree, u32 sm_count, void *out) {
    SumcheckPolys ps;
    SumcheckExpr ex;
    memset(&ps, 0, sizeof(ps));
    memset(&ex, 0, sizeof(ex));
    for (u32 p = 0; p < num_polys; ++p) ps.p[p] = (const uint4 *)polys[p];
    static const u32 FR_ONE[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    ex.num_terms = num_terms; ex.num_polys = num_polys; ex.degree = degree; ex.common = common;
    for (u32 t = 0; t < num_terms; ++t) {
        memcpy(ex.coeff[t].l, (const char *)coeffs + (size_t)t * 32, 32);
        ex.has_coeff[t] = memcmp(ex.coeff[t].l, FR_ONE, 32) != 0;
        ex.nfac[t] = (unsigned char)(offsets[t + 1] - offsets[t]);
        for (u32 j = offsets[t]; j < offsets[t + 1]; ++j) ex.fac[t][j - offsets[t]] = (unsigned char)term_polys[j];
    }
    std::vector<fe> partials((size_t)sm_count * 16 * PK_SC_MAX_DEGREE + 8);
    pk_enqueue_sumcheck_round(ps, ex, n / 2, partials.data(), out, sm_count, 0);
}
// The factored zero-check round (SumcheckExpr::common_sum): G(1..degree-1) into out; `degree` is the full round
// polynomial's degree, as for emul_sumcheck_round.
void emul_sumcheck_round_factored(const void *const *polys, u32 num_polys, u32 n, const void *coeffs, const u32 *offsets, const u32 *term_polys,
                                  u32 num_terms, int common, u32 degree, u32 sm_count, void *out) {
    SumcheckPolys ps;
    SumcheckExpr ex;
    memset(&ps, 0, sizeof(ps));
    memset(&ex, 0, sizeof(ex));
    for (u32 p = 0; p < num_polys; ++p) ps.p[p] = (const uint4 *)polys[p];
    static const u32 FR_ONE[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    ex.num_terms = num_terms; ex.num_polys = num_polys; ex.degree = degree - 1; ex.common = common; ex.common_sum = 1;
    for (u32 t = 0; t < num_terms; ++t) {
        memcpy(ex.coeff[t].l, (const char *)coeffs + (size_t)t * 32, 32);
        ex.has_coeff[t] = memcmp(ex.coeff[t].l, FR_ONE, 32) != 0;
        ex.nfac[t] = (unsigned char)(offsets[t + 1] - offsets[t]);
        for (u32 j = offsets[t]; j < offsets[t + 1]; ++j) ex.fac[t][j - offsets[t]] = (unsigned char)term_polys[j];
    }
    std::vector<fe> partials((size_t)sm_count * 16 * PK_SC_MAX_DEGREE + 8);
    pk_enqueue_sumcheck_round(ps, ex, n / 2, partials.data(), out, sm_count, 0);
}
// ---- affine tables of the sum-check compiler (poly_kernels.cuh k_fr_affine / k_fr_sparse_add)
void emul_fr_affine(const void *const *polys, const int *rotations, const void *coeffs, u32 count, u32 num_vars, const void *constant,
                    const void *id_coeff, const unsigned long long *rows, const void *values, u32 sparse_count, void *out) {
    static const u32 FR_ONE[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    AffineArgs a;
    memset(&a, 0, sizeof(a));
    a.num_vars = num_vars; a.primitive = PK_BH_PRIMITIVES[num_vars]; a.x_inv = PK_BH_X_INVS[num_vars];
    a.count = count;
    if (constant) { a.has_constant = 1; memcpy(a.constant.l, constant, 32); }
    if (id_coeff) { a.has_id = 1; memcpy(a.id_coeff.l, id_coeff, 32); }
    for (u32 i = 0; i < count; ++i) {
        a.poly[i] = (const uint4 *)polys[i];
        a.rotation[i] = rotations[i];
        memcpy(a.coeff[i].l, (const char *)coeffs + (size_t)i * 32, 32);
        if (memcmp(a.coeff[i].l, FR_ONE, 32) != 0) a.has_coeff_mask |= 1u << i;
    }
    PK_LAUNCH(k_fr_affine, dim3(3), dim3(64), 0, 0, a, (size_t)1 << num_vars, (uint4 *)out);
    if (sparse_count) PK_LAUNCH(k_fr_sparse_add, dim3(1), dim3(32), 0, 0, (uint4 *)out, rows, (const uint4 *)values, sparse_count);
}
// ---- lookup argument producers (lookup_kernels.cuh)
static void emul_fill_expr(SumcheckPolys &ps, SumcheckExpr &ex, const void *const *polys, u32 num_polys, const void *coeffs, const u32 *offsets,
                           const u32 *term_polys, u32 num_terms, int common, u32 degree) {
    memset(&ps, 0, sizeof(ps));
    memset(&ex, 0, sizeof(ex));
    for (u32 p = 0; p < num_polys; ++p) ps.p[p] = (const uint4 *)polys[p];
    static const u32 FR_ONE[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    ex.num_terms = num_terms; ex.num_polys = num_polys; ex.degree = degree; ex.common = common;
    for (u32 t = 0; t < num_terms; ++t) {
        memcpy(ex.coeff[t].l, (const char *)coeffs + (size_t)t * 32, 32);
        ex.has_coeff[t] = memcmp(ex.coeff[t].l, FR_ONE, 32) != 0;
        ex.nfac[t] = (unsigned char)(offsets[t + 1] - offsets[t]);
        for (u32 j = offsets[t]; j < offsets[t + 1]; ++j) ex.fac[t][j - offsets[t]] = (unsigned char)term_polys[j];
    }
}
void emul_expr_rows(const void *const *polys, u32 num_polys, u32 n, const void *coeffs, const u32 *offsets, const u32 *term_polys, u32 num_terms,
                    int common, void *out) {
    SumcheckPolys ps;
    SumcheckExpr ex;
    emul_fill_expr(ps, ex, polys, num_polys, coeffs, offsets, term_polys, num_terms, common, 1);
    pk_enqueue_expr_rows(ps, ex, n, out, 1, 0);
}
// returns 1 when an input value is missing from the table (the reference's Err, prover.rs:176-178)
int emul_lookup_m(const void *input, const void *table, u32 n, void *out) {
    std::vector<u32> slots(pk_lookup_slots(n)), counts(n);
    u32 missing = 0;
    pk_enqueue_lookup_m(input, table, n, slots.data(), counts.data(), &missing, out, 2, 0);
    return (int)missing;
}
void emul_lookup_h(const void *input, const void *table, const void *m, const void *gamma, u32 n, void *out) {
    pk_enqueue_lookup_h(input, table, m, gamma, n, out, 0);
}
void emul_sumcheck_fold(const void *const *polys, void *const *outs, u32 num_polys, u32 n, const void *challenge, u32 sm_count) {
    SumcheckFoldArgs a;
    memset(&a, 0, sizeof(a));
    for (u32 p = 0; p < num_polys; ++p) { a.in[p] = (const uint4 *)polys[p]; a.out[p] = (uint4 *)outs[p]; }
    pk_enqueue_sumcheck_fold(a, num_polys, challenge, n / 2, sm_count, 0);
}
}
