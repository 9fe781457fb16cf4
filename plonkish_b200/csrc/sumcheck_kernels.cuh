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
// degree * (sum of term degrees) Fr products per pair: IMAD bound for plonkish expressions.
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

// sum over the terms at the current point of every table (values in shared memory)
PK_HD fe sc_eval_terms(const SumcheckExpr &ex, const u32 *val, u32 nthreads, u32 tid) {
    fe total = fe_zero();
    for (u32 t = 0; t < ex.num_terms; ++t) {
        fe prod;
        if (ex.nfac[t] == 0) {
            prod = ex.coeff[t];
        } else {
            prod = sc_load(val, ex.fac[t][0], nthreads, tid);
            for (u32 j = 1; j < ex.nfac[t]; ++j) prod = fr_mul(prod, sc_load(val, ex.fac[t][j], nthreads, tid));
            if (ex.has_coeff[t]) prod = fr_mul(prod, ex.coeff[t]);
        }
        total = fr_add(total, prod);
    }
    if (ex.common >= 0) total = fr_mul(total, sc_load(val, (u32)ex.common, nthreads, tid));
    return total;
}

// partials[block][x-1] = sum over the block's pairs b of expr(tables at (.., X = x, b)), x = 1..degree.
// Dynamic shared memory: 2 * num_polys * 8 * blockDim words (value and step of every table per thread).
__global__ void __launch_bounds__(128) k_sumcheck_round(SumcheckPolys polys, SumcheckExpr ex, u32 size, uint4 *__restrict__ partials) {
    PK_DYN_SMEM(u32, smem);
    const u32 nt = blockDim.x, tid = threadIdx.x;
    u32 *val = smem, *step = smem + (size_t)ex.num_polys * 8 * nt;
    fe acc[PK_SC_MAX_DEGREE];
#pragma unroll
    for (int x = 0; x < PK_SC_MAX_DEGREE; ++x) acc[x] = fe_zero();
    for (u32 b = blockIdx.x * nt + tid; b < size; b += gridDim.x * nt) {
        for (u32 p = 0; p < ex.num_polys; ++p) {
            const uint4 *src = polys.p[p] + 4 * (size_t)b;      // e[2b] then e[2b+1]
            const fe lo = load_fe_plain(src), hi = load_fe_plain(src + 2);
            sc_store(val, p, nt, tid, hi);                      // X = 1
            sc_store(step, p, nt, tid, fr_sub(hi, lo));
        }
#pragma unroll
        for (int x = 0; x < PK_SC_MAX_DEGREE; ++x) {
            if ((u32)x < ex.degree) {
                acc[x] = fr_add(acc[x], sc_eval_terms(ex, val, nt, tid));
                if ((u32)x + 1 < ex.degree) {
                    for (u32 p = 0; p < ex.num_polys; ++p)     // X -> X + 1
                        sc_store(val, p, nt, tid, fr_add(sc_load(val, p, nt, tid), sc_load(step, p, nt, tid)));
                }
            }
        }
    }
    // block sum through shared memory (reusing the value slots: one fe per thread, tree over threads)
    __syncthreads();
#pragma unroll
    for (int x = 0; x < PK_SC_MAX_DEGREE; ++x) {
        if ((u32)x >= ex.degree) continue;
        sc_store(val, 0, nt, tid, acc[x]);
        __syncthreads();
        for (u32 s = nt >> 1; s >= 1; s >>= 1) {
            if (tid < s) sc_store(val, 0, nt, tid, fr_add(sc_load(val, 0, nt, tid), sc_load(val, 0, nt, tid + s)));
            __syncthreads();
        }
        if (tid == 0) store_fe(partials + 2 * ((size_t)blockIdx.x * ex.degree + x), sc_load(val, 0, nt, 0));
        __syncthreads();
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

// Threads per block such that value + step slots of every table fit in shared memory (<= 200 KB).
inline u32 pk_sumcheck_block(u32 num_polys) {
    const u32 fit = (200u * 1024u) / (64u * (num_polys ? num_polys : 1));
    u32 t = 128;  // a power of two: the block sum is a binary tree
    while (t > 32 && t > fit) t >>= 1;
    return t;
}
inline size_t pk_sumcheck_smem(u32 num_polys, u32 block) { return (size_t)64 * num_polys * block; }

// One round over tables of 2 * size evaluations each: degree values into d_out (X = 1..degree).
// partials: max_blocks * degree field elements of scratch.
inline u32 pk_sumcheck_grid(u32 size, u32 block, u32 sm_count) {
    u32 blocks = (size + block - 1) / block;
    const u32 cap = sm_count * 2;
    if (blocks > cap) blocks = cap;
    return blocks ? blocks : 1;
}
inline void pk_enqueue_sumcheck_round(const SumcheckPolys &polys, const SumcheckExpr &ex, u32 size, void *partials, void *d_out, u32 sm_count,
                                      pk_stream_t stream) {
    const u32 block = pk_sumcheck_block(ex.num_polys);
    const u32 blocks = pk_sumcheck_grid(size, block, sm_count);
    const size_t smem = pk_sumcheck_smem(ex.num_polys, block);
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
