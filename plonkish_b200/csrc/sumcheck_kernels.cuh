// Sum-check prover rounds on the device: the round polynomial of
// ClassicSumCheck<EvaluationsProver> and the variable fixing between rounds.
//
//   prove loop        /root/reference/plonkish_backend/src/piop/sum_check/classic.rs:208-240
//   round message     piop/sum_check/classic/eval.rs:101-131 (evaluations at X = 1..degree over the pairs
//                     (2b, 2b+1), eval.rs:236-243, 268-287; evals[0] = sum - evals[1], eval.rs:128)
//   next_round        classic.rs:90-141 -> MultilinearPolynomial::fix_var, poly/multilinear.rs:179-189, 599-618
//                     (out[b] = (e[2b+1] - e[2b]) * x + e[2b])
//
// The expression is handed over flattened: sum_t coeff_t * prod_j poly[fac_t,j], optionally times one common
// factor polynomial (the eq(x, y) of a zero check).  Every polynomial is an explicit table of evaluations
// over the boolean hypercube — the caller materialises eq_xy / identity / Lagrange tables and rotated copies,
// which the reference keeps implicit (classic.rs:40-75, 104-126); the round values are the same field elements.
//
// k_sumcheck_round is one pass over all tables (HBM: 64 B per polynomial and pair) with
// degree * (sum of term degrees - terms + 1) Fr products per pair: IMAD bound for plonkish expressions.
// Also compiled by g++ against tests/emul/cuda_emul.h (PLONKISH_EMUL) for the CPU suite.
#pragma once
#include "poly_kernels.cuh"

namespace pk {

#define PK_SC_MAX_POLYS 32
#define PK_SC_MAX_TERMS 32
#define PK_SC_MAX_FACTORS 8
#define PK_SC_MAX_DEGREE 8

struct SumcheckExpr {
    fe coeff[PK_SC_MAX_TERMS];                             // Montgomery Fr
    unsigned char has_coeff[PK_SC_MAX_TERMS];              // 0: coefficient is one (no multiplication)
    unsigned char nfac[PK_SC_MAX_TERMS];                   // factors of the term (0: the term is its coefficient)
    unsigned char fac[PK_SC_MAX_TERMS][PK_SC_MAX_FACTORS];  // polynomial indices
    u32 num_terms, num_polys, degree;
    int common;  // polynomial multiplying the whole sum, or -1
};
struct SumcheckPolys {
    const uint4 *p[PK_SC_MAX_POLYS];
};
struct SumcheckFoldArgs {
    const uint4 *in[PK_SC_MAX_POLYS];
    uint4 *out[PK_SC_MAX_POLYS];
};

// shared-memory slots, limb-major so that a warp reads 32 consecutive words: slot(p, limb, tid)
PK_HD fe sc_load(const u32 *base, u32 p, u32 nthreads, u32 tid) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = base[((size_t)p * 8 + i) * nthreads + tid];
    return r;
}
PK_HD void sc_store(u32 *base, u32 p, u32 nthreads, u32 tid, const fe &v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) base[((size_t)p * 8 + i) * nthreads + tid] = v.l[i];
}

// The table entries of pair b at the points X = x0 + 1 and x0 + 2: e[2b+1] + x0 * step and one step further
// (eval.rs:268-300: the point X = 1 is e[2b+1], every further point adds e[2b+1] - e[2b]).
PK_HD void sc_points(const uint4 *__restrict__ table, u32 b, u32 x0, fe &v0, fe &v1) {
    const uint4 *src = table + 4 * (size_t)b;  // e[2b] then e[2b+1]: 64 contiguous bytes
    const fe lo = load_fe_plain(src), hi = load_fe_plain(src + 2);
    const fe step = fr_sub(hi, lo);
    v0 = hi;
    for (u32 i = 0; i < x0; ++i) v0 = fr_add(v0, step);
    v1 = fr_add(v0, step);
}

// partials[block][x-1] = sum over the block's pairs b of expr(tables at (.., X = x, b)), x = 1..degree.
// Term-major: every factor is read where it is used (the 64 bytes of a pair stay in L1 across the terms and
// passes), two points of X per pass so that two independent product chains are in flight; no staging of the
// tables, so occupancy is set by registers alone.  Dynamic shared memory: degree * 8 * blockDim words (the
// running sums of every thread, limb-major).
__global__ void __launch_bounds__(128, 4) k_sumcheck_round(SumcheckPolys polys, SumcheckExpr ex, u32 size, uint4 *__restrict__ partials) {
    PK_DYN_SMEM(u32, acc);
    const u32 nt = blockDim.x, tid = threadIdx.x;
    for (u32 x = 0; x < ex.degree; ++x) sc_store(acc, x, nt, tid, fe_zero());
    for (u32 b = blockIdx.x * nt + tid; b < size; b += gridDim.x * nt) {
        for (u32 x0 = 0; x0 < ex.degree; x0 += 2) {
            const bool two = x0 + 1 < ex.degree;
            fe tot0 = fe_zero(), tot1 = fe_zero();
            for (u32 t = 0; t < ex.num_terms; ++t) {
                fe p0 = ex.coeff[t], p1 = ex.coeff[t];
                if (ex.nfac[t]) {
                    sc_points(polys.p[ex.fac[t][0]], b, x0, p0, p1);
                    for (u32 j = 1; j < ex.nfac[t]; ++j) {
                        fe v0, v1;
                        sc_points(polys.p[ex.fac[t][j]], b, x0, v0, v1);
                        p0 = fr_mul(p0, v0);
                        if (two) p1 = fr_mul(p1, v1);
                    }
                    if (ex.has_coeff[t]) {
                        p0 = fr_mul(p0, ex.coeff[t]);
                        if (two) p1 = fr_mul(p1, ex.coeff[t]);
                    }
                }
                tot0 = fr_add(tot0, p0);
                tot1 = fr_add(tot1, p1);
            }
            if (ex.common >= 0) {
                fe v0, v1;
                sc_points(polys.p[ex.common], b, x0, v0, v1);
                tot0 = fr_mul(tot0, v0);
                if (two) tot1 = fr_mul(tot1, v1);
            }
            sc_store(acc, x0, nt, tid, fr_add(sc_load(acc, x0, nt, tid), tot0));
            if (two) sc_store(acc, x0 + 1, nt, tid, fr_add(sc_load(acc, x0 + 1, nt, tid), tot1));
        }
    }
    // block sum: binary tree over the threads, one point of X at a time
    __syncthreads();
    for (u32 x = 0; x < ex.degree; ++x) {
        for (u32 s = nt >> 1; s >= 1; s >>= 1) {
            if (tid < s) sc_store(acc, x, nt, tid, fr_add(sc_load(acc, x, nt, tid), sc_load(acc, x, nt, tid + s)));
            __syncthreads();
        }
        if (tid == 0) store_fe(partials + 2 * ((size_t)blockIdx.x * ex.degree + x), sc_load(acc, x, nt, 0));
    }
}

// out[x] = sum_k partials[k][x]; one block, thread x.
__global__ void k_sumcheck_sum_partials(const uint4 *__restrict__ partials, u32 nblocks, u32 degree, uint4 *__restrict__ out) {
    const u32 x = threadIdx.x;
    if (x >= degree) return;
    fe s = fe_zero();
    for (u32 k = 0; k < nblocks; ++k) s = fr_add(s, load_fe_plain(partials + 2 * ((size_t)k * degree + x)));
    store_fe(out + 2 * (size_t)x, s);
}

// fix_var for every table in one launch (grid.y = table): out[b] = (e[2b+1] - e[2b]) * x + e[2b].
__global__ void __launch_bounds__(256) k_sumcheck_fold(SumcheckFoldArgs a, const uint4 *__restrict__ x_ptr, u32 size) {
    const fe x = load_fe_plain(x_ptr);
    const uint4 *in = a.in[blockIdx.y];
    uint4 *out = a.out[blockIdx.y];
    for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < size; b += gridDim.x * blockDim.x) {
        const fe lo = load_fe_plain(in + 4 * (size_t)b), hi = load_fe_plain(in + 4 * (size_t)b + 2);
        store_fe(out + 2 * (size_t)b, fr_add(fr_mul(fr_sub(hi, lo), x), lo));
    }
}

inline u32 pk_sumcheck_block(u32) { return 128; }  // a power of two: the block sum is a binary tree
inline size_t pk_sumcheck_smem(u32 degree, u32 block) { return (size_t)32 * degree * block; }

// One round over tables of 2 * size evaluations each: degree values into d_out (X = 1..degree).
// partials: 16 * sm_count * degree field elements of scratch (more than the largest grid).
inline u32 pk_sumcheck_grid(u32 size, u32 block, u32 sm_count) {
    u32 blocks = (size + block - 1) / block;
    const u32 cap = sm_count * 4;  // persistent: 4 blocks of 128 threads per SM (registers)
    if (blocks > cap) blocks = cap;
    return blocks ? blocks : 1;
}
inline void pk_enqueue_sumcheck_round(const SumcheckPolys &polys, const SumcheckExpr &ex, u32 size, void *partials, void *d_out, u32 sm_count,
                                      pk_stream_t stream) {
    const u32 block = pk_sumcheck_block(ex.num_polys);
    const u32 blocks = pk_sumcheck_grid(size, block, sm_count);
    const size_t smem = pk_sumcheck_smem(ex.degree, block);
    PK_SET_SMEM(k_sumcheck_round, smem);
    PK_LAUNCH(k_sumcheck_round, dim3(blocks), dim3(block), smem, stream, polys, ex, size, (uint4 *)partials);
    PK_LAUNCH(k_sumcheck_sum_partials, dim3(1), dim3(32), 0, stream, (const uint4 *)partials, blocks, ex.degree, (uint4 *)d_out);
}
inline void pk_enqueue_sumcheck_fold(const SumcheckFoldArgs &a, u32 num_polys, const void *d_challenge, u32 size, u32 sm_count, pk_stream_t stream) {
    u32 blocks = (size + 255) / 256;
    const u32 cap = sm_count * 8 / (num_polys ? num_polys : 1) + 1;
    if (blocks > cap) blocks = cap;
    PK_LAUNCH(k_sumcheck_fold, dim3(blocks, num_polys), dim3(256), 0, stream, a, (const uint4 *)d_challenge, size);
}

}  // namespace pk
