// Producers of HyperPlonk's lookup argument (logUp form) on the device — the polynomials committed at
// backend/hyperplonk.rs:227 (lookup_m_polys) and :251-252 (lookup_h_polys):
//
//   lookup_compressed_polys   /root/reference/plonkish_backend/src/backend/hyperplonk/prover.rs:50-137
//                             (an expression evaluated on every row: k_expr_rows over the compiled terms)
//   lookup_m_poly             prover.rs:145-192  (multiplicity of every table row among the inputs; a value that occurs
//                             several times in the table is counted at its LAST row: the HashMap of :151 keeps the last
//                             index inserted; an input that is not in the table is an error, :169-177)
//   lookup_h_poly             prover.rs:206-250  (h = 1 / (gamma + input) - m / (gamma + table), batch inverted)
//
// Also compiled by g++ against tests/emul/cuda_emul.h (PLONKISH_EMUL) for the CPU suite.
#pragma once
#include "sumcheck_kernels.cuh"

namespace pk {

// out[b] = (sum_t coeff_t * prod_j table[fac_t,j][b]) (* table[common][b]): a compiled expression on every row.
__global__ void __launch_bounds__(256) k_expr_rows(SumcheckPolys polys, SumcheckExpr ex, size_t n, uint4 *__restrict__ out) {
    for (size_t b = blockIdx.x * (size_t)blockDim.x + threadIdx.x; b < n; b += (size_t)gridDim.x * blockDim.x) {
        fe tot = fe_zero();
        for (u32 t = 0; t < ex.num_terms; ++t) {
            fe p = ex.coeff[t];
            if (ex.nfac[t]) {
                p = load_fe_plain(polys.p[ex.fac[t][0]] + 2 * b);
                for (u32 j = 1; j < ex.nfac[t]; ++j) p = fr_mul(p, load_fe_plain(polys.p[ex.fac[t][j]] + 2 * b));
                if (ex.has_coeff[t]) p = fr_mul(p, ex.coeff[t]);
            }
            tot = fr_add(tot, p);
        }
        if (ex.common >= 0) tot = fr_mul(tot, load_fe_plain(polys.p[ex.common] + 2 * b));
        store_fe(out + 2 * b, tot);
    }
}

#define PK_LOOKUP_EMPTY 0xffffffffu
PK_HD u32 lookup_hash(const fe &v, u32 mask) {
    u32 h = 0x9e3779b9u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h ^= v.l[i] + 0x85ebca6bu + (h << 6) + (h >> 2);
        h *= 0xc2b2ae35u;
    }
    return (h ^ (h >> 15)) & mask;
}
PK_HD bool fe_equal(const fe &a, const fe &b) {
    u32 d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.l[i] ^ b.l[i];
    return d == 0;
}
PK_HD u32 lookup_cas(u32 *p, u32 expect, u32 v) { return atomicCAS(p, expect, v); }
PK_HD u32 lookup_max(u32 *p, u32 v) { return atomicMax(p, v); }
PK_HD u32 lookup_read(const u32 *p) { return *reinterpret_cast<const volatile u32 *>(p); }

// Open-addressing table over the compressed table polynomial: slots[h] = the largest row holding that value.
__global__ void __launch_bounds__(256) k_lookup_insert(const uint4 *__restrict__ table, u32 n, u32 *__restrict__ slots, u32 mask) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const fe v = load_fe_plain(table + 2 * (size_t)i);
        u32 h = lookup_hash(v, mask);
        while (true) {
            const u32 prev = lookup_cas(&slots[h], PK_LOOKUP_EMPTY, i);
            if (prev == PK_LOOKUP_EMPTY) break;
            if (fe_equal(load_fe_plain(table + 2 * (size_t)prev), v)) { lookup_max(&slots[h], i); break; }  // same value: keep the last row
            h = (h + 1) & mask;
        }
    }
}
// counts[row] += 1 for the table row every input value maps to; *missing = 1 if some input is not in the table.
__global__ void __launch_bounds__(256) k_lookup_count(const uint4 *__restrict__ input, const uint4 *__restrict__ table, u32 n, const u32 *__restrict__ slots,
                                                      u32 mask, u32 *__restrict__ counts, u32 *__restrict__ missing) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const fe v = load_fe_plain(input + 2 * (size_t)i);
        u32 h = lookup_hash(v, mask);
        while (true) {
            const u32 row = lookup_read(&slots[h]);
            if (row == PK_LOOKUP_EMPTY) { *missing = 1u; break; }
            if (fe_equal(load_fe_plain(table + 2 * (size_t)row), v)) { atomicAdd(&counts[row], 1u); break; }
            h = (h + 1) & mask;
        }
    }
}
// m[row] = F::from(count) (prover.rs:186-190).
__global__ void __launch_bounds__(256) k_lookup_m(const u32 *__restrict__ counts, u32 n, uint4 *__restrict__ out) {
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        store_fe(out + 2 * (size_t)i, counts[i] ? fr_from_u64(counts[i]) : fe_zero());
}

// h[b] = 1 / (gamma + input[b]) - m[b] / (gamma + table[b]) (prover.rs:212-243).  One inversion per strip of rows by
// Montgomery's trick over d_b = (gamma + input[b]) * (gamma + table[b]); `out` holds the prefix products in between.
// A zero denominator (probability 2^-230 per row for a transcript challenge) would zero its strip; batch_invert skips it.
__global__ void __launch_bounds__(128) k_lookup_h(const uint4 *__restrict__ input, const uint4 *__restrict__ table, const uint4 *__restrict__ m,
                                                  const uint4 *__restrict__ gamma_ptr, size_t n, u32 strip, uint4 *__restrict__ out) {
    const size_t first = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * strip;
    if (first >= n) return;
    const u32 cnt = (u32)((first + strip <= n) ? strip : n - first);
    const fe gamma = load_fe_plain(gamma_ptr);
    fe run;
    run.l[0] = 0x4ffffffbu; run.l[1] = 0xac96341cu; run.l[2] = 0x9f60cd29u; run.l[3] = 0x36fc7695u;
    run.l[4] = 0x7879462eu; run.l[5] = 0x666ea36fu; run.l[6] = 0x9a07df2fu; run.l[7] = 0x0e0a77c1u;
    for (u32 j = 0; j < cnt; ++j) {
        const fe a = fr_add(gamma, load_fe_plain(input + 2 * (first + j))), b = fr_add(gamma, load_fe_plain(table + 2 * (first + j)));
        store_fe(out + 2 * (first + j), run);
        run = fr_mul(run, fr_mul(a, b));
    }
    fe inv = fr_inv(run);
    for (u32 j = cnt; j-- > 0;) {
        const fe a = fr_add(gamma, load_fe_plain(input + 2 * (first + j))), b = fr_add(gamma, load_fe_plain(table + 2 * (first + j)));
        const fe inv_d = fr_mul(inv, load_fe_plain(out + 2 * (first + j)));   // 1 / (a * b)
        inv = fr_mul(inv, fr_mul(a, b));
        const fe inv_a = fr_mul(inv_d, b), inv_b = fr_mul(inv_d, a);
        store_fe(out + 2 * (first + j), fr_sub(inv_a, fr_mul(inv_b, load_fe_plain(m + 2 * (first + j)))));
    }
}

inline void pk_enqueue_expr_rows(const SumcheckPolys &polys, const SumcheckExpr &ex, size_t n, void *out, u32 sm_count, pk_stream_t stream) {
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)sm_count * 8) blocks = (size_t)sm_count * 8;
    PK_LAUNCH(k_expr_rows, dim3((unsigned)blocks), dim3(256), 0, stream, polys, ex, n, (uint4 *)out);
}
// slots: 2 * n words (0xff-filled here), counts: n words (zeroed here), missing: one word (zeroed here).
inline void pk_enqueue_lookup_m(const void *input, const void *table, u32 n, u32 *slots, u32 *counts, u32 *missing, void *out, u32 sm_count,
                                pk_stream_t stream) {
    u32 cap = 2;
    while (cap < 2 * n) cap <<= 1;
    u32 blocks = (n + 255) / 256;
    if (blocks > sm_count * 8) blocks = sm_count * 8;
#ifdef PLONKISH_EMUL
    memset(slots, 0xff, sizeof(u32) * cap); memset(counts, 0, sizeof(u32) * n); *missing = 0;
#else
    cudaMemsetAsync(slots, 0xff, sizeof(u32) * cap, stream);
    cudaMemsetAsync(counts, 0, sizeof(u32) * n, stream);
    cudaMemsetAsync(missing, 0, sizeof(u32), stream);
#endif
    PK_LAUNCH(k_lookup_insert, dim3(blocks), dim3(256), 0, stream, (const uint4 *)table, n, slots, cap - 1);
    PK_LAUNCH(k_lookup_count, dim3(blocks), dim3(256), 0, stream, (const uint4 *)input, (const uint4 *)table, n, (const u32 *)slots, cap - 1, counts, missing);
    PK_LAUNCH(k_lookup_m, dim3(blocks), dim3(256), 0, stream, (const u32 *)counts, n, (uint4 *)out);
}
inline size_t pk_lookup_slots(size_t n) {
    size_t cap = 2;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}
inline void pk_enqueue_lookup_h(const void *input, const void *table, const void *m, const void *d_gamma, size_t n, void *out, pk_stream_t stream) {
    const u32 strip = pk_perm_strip(n);
    const size_t threads = (n + strip - 1) / strip;
    PK_LAUNCH(k_lookup_h, dim3((unsigned)((threads + 127) / 128)), dim3(128), 0, stream, (const uint4 *)input, (const uint4 *)table, (const uint4 *)m,
              (const uint4 *)d_gamma, n, strip, (uint4 *)out);
}

}  // namespace pk
