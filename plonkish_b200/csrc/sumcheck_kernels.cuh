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

#define PK_SC_MAX_POLYS 48
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
    // 0: the common factor walks through X like every other factor.  1: the factored zero-check round — the common
    // factor is eq(x, y) (times any constant), whose pair is (S (1 - y_r), S y_r) with S = e[2b] + e[2b+1] the eq value
    // over the remaining variables, so h(X) = (1 - y_r + X (2 y_r - 1)) * G(X) with G(X) = sum_b S_b * expr_b(X) of one
    // degree less: the kernel multiplies by S (no walk) at X = 1..degree, where `degree` is then deg G = deg h - 1, and
    // the caller rebuilds the same message h(0..deg h) from G and the running sum (plonkish_b200/sumcheck.py).
    int common_sum;
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

// Plain 256-bit addition without reduction: the values a + x * step (x <= 4, both below r) stay below 5r < 2^256 and are
// only ever used as the MULTIPLIER operand of mont_mul (read limb by limb), whose bounds depend on the multiplicand
// alone: the running value stays below a + r and the product below a * 5r / 2^256 + r < 1.95 r for a < r, so the one
// conditional subtraction at the end of mont_mul still lands in [0, r).
PK_HD fe fe_add_plain(const fe &a, const fe &b) {
    fe r;
    add8(r.l, a.l, b.l);
    return r;
}

// partials[block][x-1] = sum over the block's pairs b of expr(tables at (.., X = x, b)), x = 1..D.
// Term-major, factor by factor: a factor's pair (e[2b], e[2b+1]) is read once per term and walked through all D points
// (eval.rs:268-300: X = 1 is e[2b+1], every further point adds e[2b+1] - e[2b]) while the term's D running products —
// D independent multiplication chains — stay in registers; the first factor of a term walks in reduced form (it becomes
// the multiplicand), every other factor by plain additions (fe_add_plain) while D <= 5.  The sums over the terms of one
// pair are multiplied by the common factor's walk and added to the thread's running sums in shared memory (limb-major,
// D * 8 * blockDim words).
template <int D>
__global__ void __launch_bounds__(128, (D <= 5 ? 3 : 2)) k_sumcheck_round_t(SumcheckPolys polys, SumcheckExpr ex, u32 size, uint4 *__restrict__ partials) {
    PK_DYN_SMEM(u32, acc);
    constexpr bool PLAIN = D <= 5;
    const u32 nt = blockDim.x, tid = threadIdx.x;
    for (u32 x = 0; x < (u32)D; ++x) sc_store(acc, x, nt, tid, fe_zero());
    for (u32 b = blockIdx.x * nt + tid; b < size; b += gridDim.x * nt) {
        fe tot[D];
#pragma unroll
        for (int x = 0; x < D; ++x) tot[x] = fe_zero();
        for (u32 t = 0; t < ex.num_terms; ++t) {
            fe prod[D];
            if (ex.nfac[t] == 0) {
#pragma unroll
                for (int x = 0; x < D; ++x) tot[x] = fr_add(tot[x], ex.coeff[t]);
                continue;
            }
            {
                const uint4 *src = polys.p[ex.fac[t][0]] + 4 * (size_t)b;
                const fe lo = load_fe_plain(src), hi = load_fe_plain(src + 2);
                const fe step = fr_sub(hi, lo);
                prod[0] = hi;
#pragma unroll
                for (int x = 1; x < D; ++x) prod[x] = fr_add(prod[x - 1], step);
            }
            for (u32 j = 1; j < ex.nfac[t]; ++j) {
                const uint4 *src = polys.p[ex.fac[t][j]] + 4 * (size_t)b;
                const fe lo = load_fe_plain(src), hi = load_fe_plain(src + 2);
                const fe step = fr_sub(hi, lo);
                fe v = hi;
#pragma unroll
                for (int x = 0; x < D; ++x) {
                    if (x) v = PLAIN ? fe_add_plain(v, step) : fr_add(v, step);
                    prod[x] = fr_mul(prod[x], v);
                }
            }
            if (ex.has_coeff[t]) {
#pragma unroll
                for (int x = 0; x < D; ++x) prod[x] = fr_mul(prod[x], ex.coeff[t]);
            }
#pragma unroll
            for (int x = 0; x < D; ++x) tot[x] = fr_add(tot[x], prod[x]);
        }
        if (ex.common >= 0) {
            const uint4 *src = polys.p[ex.common] + 4 * (size_t)b;
            const fe lo = load_fe_plain(src), hi = load_fe_plain(src + 2);
            if (ex.common_sum) {
                const fe v = PLAIN ? fe_add_plain(lo, hi) : fr_add(lo, hi);  // < 2r: a multiplier, like the walked values
#pragma unroll
                for (int x = 0; x < D; ++x) tot[x] = fr_mul(tot[x], v);
            } else {
                const fe step = fr_sub(hi, lo);
                fe v = hi;
#pragma unroll
                for (int x = 0; x < D; ++x) {
                    if (x) v = PLAIN ? fe_add_plain(v, step) : fr_add(v, step);
                    tot[x] = fr_mul(tot[x], v);
                }
            }
        }
#pragma unroll
        for (int x = 0; x < D; ++x) sc_store(acc, x, nt, tid, fr_add(sc_load(acc, x, nt, tid), tot[x]));
    }
    // block sum: binary tree over the threads, one point of X at a time
    __syncthreads();
    for (u32 x = 0; x < (u32)D; ++x) {
        for (u32 s = nt >> 1; s >= 1; s >>= 1) {
            if (tid < s) sc_store(acc, x, nt, tid, fr_add(sc_load(acc, x, nt, tid), sc_load(acc, x, nt, tid + s)));
            __syncthreads();
        }
        if (tid == 0) store_fe(partials + 2 * ((size_t)blockIdx.x * D + x), sc_load(acc, x, nt, 0));
    }
}

// out[x] = sum_k partials[k][x]: block x, 128 threads stride over the blocks' partial sums, then a binary tree in shared
// memory (a single thread per point took 65 us for 444 partials, 3 ms over the 48 rounds of a k = 24 proof).
__global__ void __launch_bounds__(128) k_sumcheck_sum_partials(const uint4 *__restrict__ partials, u32 nblocks, u32 degree, uint4 *__restrict__ out) {
    __shared__ fe part[128];
    const u32 x = blockIdx.x, tid = threadIdx.x;
    fe s = fe_zero();
    for (u32 k = tid; k < nblocks; k += blockDim.x) s = fr_add(s, load_fe_plain(partials + 2 * ((size_t)k * degree + x)));
    part[tid] = s;
    __syncthreads();
    for (u32 w = blockDim.x >> 1; w >= 1; w >>= 1) {
        if (tid < w) part[tid] = fr_add(part[tid], part[tid + w]);
        __syncthreads();
    }
    if (tid == 0) store_fe(out + 2 * (size_t)x, part[0]);
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
inline u32 pk_sumcheck_blocks_per_sm(u32 degree) { return degree <= 5 ? 3u : 2u; }  // the launch bounds of k_sumcheck_round_t
inline u32 pk_sumcheck_grid(u32 size, u32 block, u32 sm_count, u32 degree) {
    u32 blocks = (size + block - 1) / block;
    const u32 cap = sm_count * pk_sumcheck_blocks_per_sm(degree);  // persistent: one wave of resident blocks
    if (blocks > cap) blocks = cap;
    return blocks ? blocks : 1;
}
inline void pk_enqueue_sumcheck_round(const SumcheckPolys &polys, const SumcheckExpr &ex, u32 size, void *partials, void *d_out, u32 sm_count,
                                      pk_stream_t stream) {
    const u32 block = pk_sumcheck_block(ex.num_polys);
    const u32 blocks = pk_sumcheck_grid(size, block, sm_count, ex.degree);
    const size_t smem = pk_sumcheck_smem(ex.degree, block);
#define PK_SC_CASE(DEG)                                                                                                    \
    case DEG:                                                                                                              \
        PK_SET_SMEM(k_sumcheck_round_t<DEG>, smem);                                                                        \
        PK_LAUNCH(k_sumcheck_round_t<DEG>, dim3(blocks), dim3(block), smem, stream, polys, ex, size, (uint4 *)partials);   \
        break;
    switch (ex.degree) {
        PK_SC_CASE(1) PK_SC_CASE(2) PK_SC_CASE(3) PK_SC_CASE(4) PK_SC_CASE(5) PK_SC_CASE(6) PK_SC_CASE(7) PK_SC_CASE(8)
        default: break;  // plonkish_cuda_sumcheck_new refuses degrees outside 1..PK_SC_MAX_DEGREE
    }
#undef PK_SC_CASE
    PK_LAUNCH(k_sumcheck_sum_partials, dim3(ex.degree), dim3(128), 0, stream, (const uint4 *)partials, blocks, ex.degree, (uint4 *)d_out);
}
inline void pk_enqueue_sumcheck_fold(const SumcheckFoldArgs &a, u32 num_polys, const void *d_challenge, u32 size, u32 sm_count, pk_stream_t stream) {
    u32 blocks = (size + 255) / 256;
    const u32 cap = sm_count * 8 / (num_polys ? num_polys : 1) + 1;
    if (blocks > cap) blocks = cap;
    PK_LAUNCH(k_sumcheck_fold, dim3(blocks, num_polys), dim3(256), 0, stream, a, (const uint4 *)d_challenge, size);
}

}  // namespace pk
