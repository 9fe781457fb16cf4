// Kernels for the callers either side of the MSM: the scalar
// producers of MultilinearKzg::open and the SRS construction of MultilinearKzg::setup.
//
//   quotients        /root/reference/plonkish_backend/src/pcs/multilinear.rs:72-107
//   g_prime merge    pcs/multilinear.rs:203-213
//   eq tables        pcs/multilinear/kzg.rs:174-193
//   window_table     util/arithmetic/msm.rs:16-31
//   fixed_base_msm   util/arithmetic/msm.rs:50-81, then batch_normalize (kzg.rs:204-207)
//
// The Fr kernels are one multiplication per 64-96 bytes moved: HBM bound, 32-byte elements
// moved as two 16-byte words.  The fixed-base kernel is the accumulate loop's mixed addition
// against a small table that stays in L2: IMAD bound like k_accumulate.
//
// Also compiled by g++ against tests/emul/cuda_emul.h (PLONKISH_EMUL) for the CPU suite.
#pragma once
#include "msm_kernels.cuh"

namespace pk {

PK_HD fe fr_mul(const fe &a, const fe &b) { return mont_mul<FrMod>(a, b); }
PK_HD fe fr_add(const fe &a, const fe &b) { return mod_add<FrMod>(a, b); }
PK_HD fe fr_sub(const fe &a, const fe &b) { return mod_sub<FrMod>(a, b); }

// plain (coherent) 32-byte load: for buffers another launch of the same stream wrote
PK_HD fe load_fe_plain(const uint4 *p) {
    const uint4 a = p[0], b = p[1];
    fe r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// ------------------------------------------------------------------ quotients
// One level of `quotients` (multilinear.rs:85-97): q[j] = hi[j] - lo[j],
// rem_out[j] = lo[j] + q[j] * x with lo = rem_in[0..half), hi = rem_in[half..2*half).
__global__ void __launch_bounds__(256) k_quotient_fold(const uint4 *__restrict__ rem_in, uint4 *__restrict__ rem_out, uint4 *__restrict__ q,
                                                       const uint4 *__restrict__ x_ptr, u32 half) {
    const fe x = load_fe_plain(x_ptr);
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < half; j += gridDim.x * blockDim.x) {
        const fe lo = load_fe_plain(rem_in + 2 * (size_t)j);
        const fe hi = load_fe_plain(rem_in + 2 * ((size_t)half + j));
        const fe d = fr_sub(hi, lo);
        store_fe(q + 2 * (size_t)j, d);
        store_fe(rem_out + 2 * (size_t)j, fr_add(lo, fr_mul(d, x)));
    }
}

// The last `levels` (<= 10) levels in one block: the remainder lives in shared memory.
// q_base is the quotient buffer (q_i at element offset 2^i); point[i] = x_i.
#define PK_QTAIL_MAX_LEVELS 10
__global__ void __launch_bounds__(512) k_quotient_tail(const uint4 *__restrict__ rem_in, uint4 *__restrict__ q_base, const uint4 *__restrict__ point,
                                                       u32 levels, uint4 *__restrict__ eval_out) {
    __shared__ fe rem[1u << PK_QTAIL_MAX_LEVELS];
    const u32 n = 1u << levels;
    for (u32 j = threadIdx.x; j < n; j += blockDim.x) rem[j] = load_fe_plain(rem_in + 2 * (size_t)j);
    __syncthreads();
    for (u32 i = levels; i-- > 0;) {
        const u32 half = 1u << i;
        const fe x = load_fe_plain(point + 2 * (size_t)i);
        // half <= 512 = blockDim: one element per thread
        fe lo, d;
        const bool on = threadIdx.x < half;
        if (on) {
            lo = rem[threadIdx.x];
            d = fr_sub(rem[half + threadIdx.x], lo);
            store_fe(q_base + 2 * ((size_t)half + threadIdx.x), d);
        }
        __syncthreads();
        if (on) rem[threadIdx.x] = fr_add(lo, fr_mul(d, x));
        __syncthreads();
    }
    if (threadIdx.x == 0) store_fe(eval_out, rem[0]);
}

// Host side of `quotients`: poly (2^k scalars) -> q buffer (2^k entries, q_i at offset 2^i),
// f(point) -> eval_out.  rem_a holds 2^(k-1) scalars, rem_b 2^(k-2).
inline void pk_enqueue_quotients(const void *poly, u32 k, const void *d_point, void *q, void *rem_a, void *rem_b, void *eval_out,
                                 u32 sm_count, pk_stream_t stream) {
    const uint4 *src = (const uint4 *)poly;
    uint4 *bufs[2] = {(uint4 *)rem_a, (uint4 *)rem_b};
    int which = 0;
    u32 level = k;
    while (level > PK_QTAIL_MAX_LEVELS) {
        const u32 i = level - 1, half = 1u << i;
        u32 blocks = (half + 255) / 256;
        const u32 cap = sm_count * 8;
        if (blocks > cap) blocks = cap;
        PK_LAUNCH(k_quotient_fold, dim3(blocks), dim3(256), 0, stream, src, bufs[which], (uint4 *)q + 2 * (size_t)half,
                  (const uint4 *)d_point + 2 * (size_t)i, half);
        src = bufs[which];
        which ^= 1;
        level = i;
    }
    PK_LAUNCH(k_quotient_tail, dim3(1), dim3(512), 0, stream, src, (uint4 *)q, (const uint4 *)d_point, level, (uint4 *)eval_out);
}

// -------------------------------------------------------------- g_prime merge
#define PK_LINCOMB_MAX 12
struct LincombArgs {
    const uint4 *poly[PK_LINCOMB_MAX];
    fe coeff[PK_LINCOMB_MAX];
    unsigned long long len[PK_LINCOMB_MAX];  // polynomial i holds len[i] values; past them it counts as zero (univariate sums of mixed degrees)
    u32 count;
    u32 accumulate;  // 1: add to what `out` already holds (more than PK_LINCOMB_MAX terms)
};
// out[j] (+)= sum_i coeff[i] * poly[i][j]   (multilinear.rs:208-213; `f += (scalar, q)` of poly/univariate.rs for mixed lengths)
__global__ void __launch_bounds__(256) k_fr_lincomb(LincombArgs a, size_t n, uint4 *__restrict__ out) {
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        fe acc = a.accumulate ? load_fe_plain(out + 2 * j) : fe_zero();
        for (u32 i = 0; i < a.count; ++i) {
            if (j < a.len[i]) acc = fr_add(acc, fr_mul(a.coeff[i], load_fe(a.poly[i] + 2 * j)));
        }
        store_fe(out + 2 * j, acc);
    }
}

// ------------------------------------------------------------------ Gemini folds
// merge_into(.., x_i, 1, 0) (poly/multilinear.rs:599-618) as Gemini::open applies it (pcs/multilinear/gemini.rs:101-108):
// out[j] = (in[2j + 1] - in[2j]) * x + in[2j] for j < half.
__global__ void __launch_bounds__(256) k_fr_fold_pairs(const uint4 *__restrict__ in, uint4 *__restrict__ out, const uint4 *__restrict__ x_ptr, size_t half) {
    const fe x = load_fe_plain(x_ptr);
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < half; j += (size_t)gridDim.x * blockDim.x) {
        const fe e0 = load_fe_plain(in + 4 * j);
        const fe e1 = load_fe_plain(in + 4 * j + 2);
        store_fe(out + 2 * j, fr_add(fr_mul(fr_sub(e1, e0), x), e0));
    }
}
// fs[1..num_vars) of Gemini::open, packed like the quotients: f_i (2^(num_vars - i) values) at element offset
// 2^(num_vars - i) of `out` (2^num_vars scalars; elements 0 and 1 are not written).  point[i - 1] folds f_(i-1) into f_i.
inline void pk_enqueue_gemini_folds(const void *poly, u32 num_vars, const void *d_point, void *out, u32 sm_count, pk_stream_t stream) {
    const uint4 *src = (const uint4 *)poly;
    for (u32 i = 1; i < num_vars; ++i) {
        const size_t half = (size_t)1 << (num_vars - i);
        uint4 *dst = (uint4 *)out + 2 * half;
        size_t blocks = (half + 255) / 256;
        const size_t cap = (size_t)sm_count * 8;
        if (blocks > cap) blocks = cap;
        PK_LAUNCH(k_fr_fold_pairs, dim3((unsigned)blocks), dim3(256), 0, stream, src, dst, (const uint4 *)d_point + 2 * (size_t)(i - 1), half);
        src = dst;
    }
}

// ------------------------------------------------------------------ eq tables
// kzg.rs:183-191: hi[j] = s * last[j], lo[j] = last[j] - hi[j].
__global__ void __launch_bounds__(256) k_eq_expand(const uint4 *__restrict__ last, uint4 *__restrict__ lo, uint4 *__restrict__ hi,
                                                   const uint4 *__restrict__ s_ptr, u32 len) {
    const fe s = load_fe_plain(s_ptr);
    for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) {
        const fe e = load_fe_plain(last + 2 * (size_t)j);
        const fe h = fr_mul(s, e);
        store_fe(hi + 2 * (size_t)j, h);
        store_fe(lo + 2 * (size_t)j, fr_sub(e, h));
    }
}
__global__ void k_fr_set_one(uint4 *out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        fe one;  // R mod r: Montgomery form of 1
        one.l[0] = 0x4ffffffbu; one.l[1] = 0xac96341cu; one.l[2] = 0x9f60cd29u; one.l[3] = 0x36fc7695u;
        one.l[4] = 0x7879462eu; one.l[5] = 0x666ea36fu; one.l[6] = 0x9a07df2fu; one.l[7] = 0x0e0a77c1u;
        store_fe(out, one);
    }
}
// eqs as scalars: slice k (2^k values) at element offset 2^k - 1 of `out` (kzg.rs:199 order).
inline void pk_enqueue_eq_scalars(const void *d_ss, u32 num_vars, void *out, u32 sm_count, pk_stream_t stream) {
    uint4 *o = (uint4 *)out;
    PK_LAUNCH(k_fr_set_one, dim3(1), dim3(32), 0, stream, o);
    for (u32 k = 0; k < num_vars; ++k) {
        const u32 len = 1u << k;
        u32 blocks = (len + 255) / 256;
        const u32 cap = sm_count * 8;
        if (blocks > cap) blocks = cap;
        PK_LAUNCH(k_eq_expand, dim3(blocks), dim3(256), 0, stream, (const uint4 *)(o + 2 * ((size_t)len - 1)), o + 2 * ((size_t)2 * len - 1),
                  o + 2 * ((size_t)3 * len - 1), (const uint4 *)d_ss + 2 * (size_t)k, len);
    }
}

// ------------------------------------------------------------------ powers of s
// powers(s).take(n) (pcs/univariate/kzg.rs:180): out[i] = s^i.  Thread t starts its chunk of PK_POW_CHUNK consecutive
// exponents at s^(t * chunk) (square-and-multiply) and walks it with one product per element.
#define PK_POW_CHUNK 256
__global__ void __launch_bounds__(128) k_fr_powers(const uint4 *__restrict__ s_ptr, size_t n, uint4 *__restrict__ out) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t first = t * PK_POW_CHUNK;
    if (first >= n) return;
    const fe s = load_fe_plain(s_ptr);
    fe one;  // R mod r
    one.l[0] = 0x4ffffffbu; one.l[1] = 0xac96341cu; one.l[2] = 0x9f60cd29u; one.l[3] = 0x36fc7695u;
    one.l[4] = 0x7879462eu; one.l[5] = 0x666ea36fu; one.l[6] = 0x9a07df2fu; one.l[7] = 0x0e0a77c1u;
    fe cur = one;
    for (int bit = 63; bit >= 0; --bit) {  // cur = s^first
        cur = fr_mul(cur, cur);
        if ((first >> bit) & 1) cur = fr_mul(cur, s);
    }
    const size_t end = (first + PK_POW_CHUNK < n) ? first + PK_POW_CHUNK : n;
    for (size_t i = first; i < end; ++i) {
        store_fe(out + 2 * i, cur);
        cur = fr_mul(cur, s);
    }
}
inline void pk_enqueue_fr_powers(const void *d_s, size_t n, void *out, pk_stream_t stream) {
    const size_t threads = (n + PK_POW_CHUNK - 1) / PK_POW_CHUNK;
    PK_LAUNCH(k_fr_powers, dim3((unsigned)((threads + 127) / 128)), dim3(128), 0, stream, (const uint4 *)d_s, n, (uint4 *)out);
}

// ------------------------------------------------------- division by (X - z)
// UnivariatePolynomial::div_rem by the divisor (X - z) (poly/univariate.rs:144-168 as called from UnivariateKzg::open,
// pcs/univariate/kzg.rs:281-282, and, point after point, for the vanishing polynomials of batch_open, :327): with
// h[i] = c[i] + z h[i+1] (h[n] = 0) the quotient is q[i-1] = h[i] and the remainder h[0].  The recurrence is cut into
// chunks of 2^log_chunk coefficients: every thread first runs its chunk with a zero carry-in, the chunk totals obey
// the same recurrence with z^chunk (solved by recursion, five levels for 2^24 coefficients at chunk 16), and a second pass adds
// z^(distance) * carry-in.  Three products per coefficient, 32-byte accesses.
#define PK_HORNER_LOG_CHUNK 4   // chunk = 16 coefficients per thread (256 until the last part of round 2: DESIGN.md 6d has the sweep)
__global__ void __launch_bounds__(128) k_horner_local(const uint4 *__restrict__ c, size_t n, const uint4 *__restrict__ z_ptr, uint4 *__restrict__ lh,
                                                      uint4 *__restrict__ totals, u32 log_chunk) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t chunk = (size_t)1 << log_chunk;
    const size_t first = t * chunk;
    if (first >= n) return;
    const size_t end = (first + chunk < n) ? first + chunk : n;
    const fe z = load_fe_plain(z_ptr);
    fe acc = fe_zero();
    for (size_t i = end; i-- > first;) {
        acc = fr_add(load_fe_plain(c + 2 * i), fr_mul(z, acc));
        store_fe(lh + 2 * i, acc);
    }
    store_fe(totals + 2 * t, acc);
}
// h[i] = lh[i] + z^(end - i) * carry[t + 1] for chunk t (carry = the solved recurrence of the totals; carry[chunks] = 0 is
// not stored: the last chunk is final as it is).  shift = 1 writes h[i] to out[i - 1] and h[0] to rem (quotient layout).
__global__ void __launch_bounds__(128) k_horner_fix(const uint4 *__restrict__ lh, size_t n, const uint4 *__restrict__ z_ptr, const uint4 *__restrict__ carry,
                                                    size_t chunks, uint4 *__restrict__ out, int shift, uint4 *__restrict__ rem, u32 log_chunk) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t chunk = (size_t)1 << log_chunk;
    const size_t first = t * chunk;
    if (first >= n) return;
    const size_t end = (first + chunk < n) ? first + chunk : n;
    const fe z = load_fe_plain(z_ptr);
    const bool has = t + 1 < chunks;
    const fe cin = has ? load_fe_plain(carry + 2 * (t + 1)) : fe_zero();
    fe pw = z;
    for (size_t i = end; i-- > first;) {
        fe v = load_fe_plain(lh + 2 * i);
        if (has) {
            v = fr_add(v, fr_mul(pw, cin));
            pw = fr_mul(pw, z);
        }
        if (!shift) store_fe(out + 2 * i, v);
        else if (i) store_fe(out + 2 * (i - 1), v);
        else store_fe(rem, v);
    }
}
// one thread: the whole recurrence of a short array (the top of the recursion)
__global__ void k_horner_serial(const uint4 *__restrict__ c, size_t n, const uint4 *__restrict__ z_ptr, uint4 *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const fe z = load_fe_plain(z_ptr);
    fe acc = fe_zero();
    for (size_t i = n; i-- > 0;) {
        acc = fr_add(load_fe_plain(c + 2 * i), fr_mul(z, acc));
        store_fe(out + 2 * i, acc);
    }
}
// zs[l + 1] = zs[l]^chunk
#define PK_HORNER_Z_SLOTS 8
__global__ void k_horner_pow(uint4 *__restrict__ zs, u32 levels, u32 log_chunk) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    fe z = load_fe_plain(zs);
    for (u32 l = 0; l < levels; ++l) {
        for (u32 k = 0; k < log_chunk; ++k) z = fr_mul(z, z);
        store_fe(zs + 2 * (size_t)(l + 1), z);
    }
}
inline size_t pk_horner_scratch_elems(size_t n, u32 log_chunk = PK_HORNER_LOG_CHUNK) {  // lh of every level + the totals
    const size_t chunk = (size_t)1 << log_chunk;
    size_t total = 8, m = n;
    while (m > 64) {
        total += m;
        m = (m + chunk - 1) / chunk;
        total += m;  // totals (the next level's input)
    }
    return total + 64 + 64;
}
// out = h (shift 0) or the quotient with the remainder in *rem (shift 1).  c has n coefficients; zs[0] = z on entry
// (PK_HORNER_Z_SLOTS element slots: 2^28 coefficients need 6 levels at chunk 16); scratch holds pk_horner_scratch_elems(n) elements.
inline void pk_enqueue_horner(const uint4 *c, size_t n, uint4 *zs, u32 level, uint4 *scratch, uint4 *out, int shift, uint4 *rem, pk_stream_t stream,
                              u32 log_chunk) {
    if (n <= 64) {
        if (!shift) {
            PK_LAUNCH(k_horner_serial, dim3(1), dim3(32), 0, stream, c, n, (const uint4 *)(zs + 2 * (size_t)level), out);
        } else {  // serial into scratch, then a one-chunk fix pass does the shifted copy
            PK_LAUNCH(k_horner_serial, dim3(1), dim3(32), 0, stream, c, n, (const uint4 *)(zs + 2 * (size_t)level), scratch);
            PK_LAUNCH(k_horner_fix, dim3(1), dim3(128), 0, stream, (const uint4 *)scratch, n, (const uint4 *)(zs + 2 * (size_t)level), (const uint4 *)scratch, (size_t)1, out, 1, rem, 6u);
        }
        return;
    }
    const size_t chunk = (size_t)1 << log_chunk;
    const size_t chunks = (n + chunk - 1) / chunk;
    uint4 *lh = scratch, *totals = scratch + 2 * n, *next = totals + 2 * chunks;
    const unsigned blocks = (unsigned)((chunks + 127) / 128);
    PK_LAUNCH(k_horner_local, dim3(blocks), dim3(128), 0, stream, c, n, (const uint4 *)(zs + 2 * (size_t)level), lh, totals, log_chunk);
    // the totals obey the same recurrence with z^chunk: solve in place of `totals` (h of the totals), using what follows as scratch
    pk_enqueue_horner(totals, chunks, zs, level + 1, next, totals, 0, nullptr, stream, log_chunk);
    PK_LAUNCH(k_horner_fix, dim3(blocks), dim3(128), 0, stream, (const uint4 *)lh, n, (const uint4 *)(zs + 2 * (size_t)level), (const uint4 *)totals, chunks, out, shift, rem, log_chunk);
}
// q (n - 1 coefficients, written to q[0 .. n-1); q[n-1] is set to zero) and rem = c mod (X - z).  z in zs[0].
inline void pk_enqueue_div_linear(const void *c, size_t n, void *zs, void *scratch, void *q, void *rem, pk_stream_t stream,
                                  u32 log_chunk = PK_HORNER_LOG_CHUNK) {
    PK_LAUNCH(k_horner_pow, dim3(1), dim3(32), 0, stream, (uint4 *)zs, (u32)(PK_HORNER_Z_SLOTS - 1), log_chunk);
    PK_MEMSET0((uint4 *)q + 2 * (n - 1), 32, stream);
    pk_enqueue_horner((const uint4 *)c, n, (uint4 *)zs, 0, (uint4 *)scratch, (uint4 *)q, 1, (uint4 *)rem, stream, log_chunk);
}

// ------------------------------------------------ permutation grand products
// permutation_z_polys (backend/hyperplonk/prover.rs:252-345), the producer of the polynomial HyperPlonk commits at
// backend/hyperplonk.rs:251-252.  Per chunk of permutation polynomials and per row b:
//     product[b] = prod_i (beta * id_i(b) + gamma + value_i(b)) / prod_i (beta * sigma_i(b) + gamma + value_i(b)),
// id_i(b) = (idx_i << num_vars) + b; then the running product over the rows in the order the reference's
// BooleanHypercube iterates them (util/arithmetic/bh.rs:118-125: 0, then 1, X, X^2, ... in GF(2^k) modulo a primitive
// polynomial), chunk after chunk within a row, and the result stored back in index order (nth_map, bh.rs:127-133).

// a^(r-2) in Fr, 0 -> 0 (Field::invert's value; batch_invert at prover.rs:282 leaves zeros alone as well).
PK_HD fe fr_inv(const fe &a) {
    u32 e[8];
    FrMod::limbs(e);
    e[0] -= 2;  // r ends in ...0001: 0xf0000001 - 2, no borrow
    fe one;     // R mod r
    one.l[0] = 0x4ffffffbu; one.l[1] = 0xac96341cu; one.l[2] = 0x9f60cd29u; one.l[3] = 0x36fc7695u;
    one.l[4] = 0x7879462eu; one.l[5] = 0x666ea36fu; one.l[6] = 0x9a07df2fu; one.l[7] = 0x0e0a77c1u;
    fe acc = one, base = a;
    for (int i = 0; i < 254; ++i) {
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fr_mul(acc, base);
        base = fr_mul(base, base);
    }
    return acc;
}
// Montgomery form of a small integer: v * R mod r = mont_mul(v, R^2 mod r).
PK_HD fe fr_from_u64(unsigned long long v) {
    fe x = fe_zero(), r2;
    x.l[0] = (u32)v; x.l[1] = (u32)(v >> 32);
    r2.l[0] = 0xae216da7u; r2.l[1] = 0x1bb8e645u; r2.l[2] = 0xe35c59e3u; r2.l[3] = 0x53fe3ab1u;
    r2.l[4] = 0x53bb8085u; r2.l[5] = 0x8c49833du; r2.l[6] = 0x7f4e44a5u; r2.l[7] = 0x0216d0b1u;
    return fr_mul(x, r2);
}

#define PK_PERM_MAX 8     // permutation polynomials per chunk (vanilla_plonk has three in one chunk)
#define PK_PERM_STRIP 32  // rows per thread at small sizes: one inversion (~370 products) per strip (Montgomery's trick)
// From 2^20 rows a thread takes 128: the inversion's share drops from 11.5 to 2.9 products per row and 2^24 rows still fill the GPU 7 times over.
inline u32 pk_perm_strip(size_t n) { return n >= ((size_t)1 << 20) ? 128u : (u32)PK_PERM_STRIP; }
struct PermArgs {
    const uint4 *value[PK_PERM_MAX];  // polys[*poly]: the witness column the permutation polynomial belongs to
    const uint4 *sigma[PK_PERM_MAX];  // the permutation polynomial
    unsigned long long id_offset[PK_PERM_MAX];  // idx << num_vars
    u32 count;
};
__global__ void __launch_bounds__(128) k_perm_products(PermArgs a, const uint4 *__restrict__ beta_gamma, size_t n, u32 strip, uint4 *__restrict__ den_cache,
                                                       uint4 *__restrict__ product) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t first = t * strip;
    if (first >= n) return;
    const u32 cnt = (u32)((first + strip <= n) ? strip : n - first);
    const fe beta = load_fe_plain(beta_gamma), gamma = load_fe_plain(beta_gamma + 2);
    fe one;
    one.l[0] = 0x4ffffffbu; one.l[1] = 0xac96341cu; one.l[2] = 0x9f60cd29u; one.l[3] = 0x36fc7695u;
    one.l[4] = 0x7879462eu; one.l[5] = 0x666ea36fu; one.l[6] = 0x9a07df2fu; one.l[7] = 0x0e0a77c1u;
    // pass 1: prefix products of the denominators (left in `product`), the denominators themselves in den_cache
    fe run = one;
    for (u32 j = 0; j < cnt; ++j) {
        fe den = one;
        for (u32 i = 0; i < a.count; ++i)
            den = fr_mul(den, fr_add(fr_add(fr_mul(beta, load_fe(a.sigma[i] + 2 * (first + j))), gamma), load_fe(a.value[i] + 2 * (first + j))));
        store_fe(product + 2 * (first + j), run);      // prefix before row j
        store_fe(den_cache + 2 * (first + j), den);
        run = fr_mul(run, den);
    }
    fe inv = fr_inv(run);                               // a zero denominator zeroes the strip, as batch_invert would skip it: not reachable for a valid beta, gamma
    // pass 2, backwards: 1 / den_j = inv * prefix_j, then the numerator
    fe beta_id[PK_PERM_MAX];
    for (u32 i = 0; i < a.count; ++i) beta_id[i] = fr_mul(beta, fr_from_u64(a.id_offset[i] + first + cnt - 1));
    for (u32 j = cnt; j-- > 0;) {
        fe num = one;
        for (u32 i = 0; i < a.count; ++i) {
            num = fr_mul(num, fr_add(fr_add(beta_id[i], gamma), load_fe(a.value[i] + 2 * (first + j))));
            beta_id[i] = fr_sub(beta_id[i], beta);
        }
        const fe inv_den = fr_mul(inv, load_fe_plain(product + 2 * (first + j)));
        inv = fr_mul(inv, load_fe_plain(den_cache + 2 * (first + j)));
        store_fe(product + 2 * (first + j), fr_mul(num, inv_den));
    }
}

// X^e in GF(2)[X] / primitive (degree k), e >= 0: the (e + 1)-th element of BooleanHypercube::iter after the leading 0.
PK_HD u32 bh_pow_x(unsigned long long e, u32 k, u32 primitive) {
    if (k == 0) return 1u;
    auto mul = [&](u32 x, u32 y) {
        unsigned long long acc = 0;
        for (u32 i = 0; i < k; ++i)
            if ((y >> i) & 1u) acc ^= (unsigned long long)x << i;
        for (int i = (int)(2 * k) - 2; i >= (int)k; --i)
            if ((acc >> i) & 1ull) acc ^= (unsigned long long)primitive << (i - k);
        return (u32)acc;
    };
    u32 result = 1u, base = (k == 1) ? (2u ^ primitive) & 1u : 2u;  // X itself (reduced when k = 1: X = 1 mod X + 1)
    while (e) {
        if (e & 1ull) result = mul(result, base);
        base = mul(base, base);
        e >>= 1;
    }
    return result;
}
PK_HD u32 bh_next(u32 b, u32 k, u32 primitive) {  // bh.rs:141-146
    b <<= 1;
    b ^= (b >> k) * primitive;
    return b;
}
// The factor sequence f_j = products[j % nc][bh(1 + j / nc)], j < m, gathered in scan order, with the running product of
// every PK_SCAN_STRIP consecutive factors taken on the way: seq[j] = f_first * ... * f_j within the strip, totals[t] = the
// strip's product.
#define PK_SCAN_STRIP 256
struct PermSeq {
    const uint4 *product[PK_PERM_MAX];  // per chunk
    u32 nc, k, primitive;
};
__global__ void __launch_bounds__(128) k_perm_gather_scan(PermSeq q, size_t m, uint4 *__restrict__ seq, uint4 *__restrict__ totals) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t first = t * PK_SCAN_STRIP;
    if (first >= m) return;
    const size_t end = (first + PK_SCAN_STRIP < m) ? first + PK_SCAN_STRIP : m;
    size_t nth = 1 + first / q.nc;       // row number in hypercube order
    u32 c = (u32)(first % q.nc);
    u32 b = bh_pow_x(nth - 1, q.k, q.primitive);
    fe run = load_fe_plain(q.product[c] + 2 * (size_t)b);
    store_fe(seq + 2 * first, run);
    for (size_t j = first + 1; j < end; ++j) {
        if (++c == q.nc) { c = 0; b = bh_next(b, q.k, q.primitive); }
        run = fr_mul(run, load_fe_plain(q.product[c] + 2 * (size_t)b));
        store_fe(seq + 2 * j, run);
    }
    store_fe(totals + 2 * t, run);
}
// in-place inclusive prefix product of a short array (the top of the recursion)
__global__ void k_prodscan_serial(uint4 *__restrict__ v, size_t n) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    fe run = load_fe_plain(v);
    for (size_t i = 1; i < n; ++i) {
        run = fr_mul(run, load_fe_plain(v + 2 * i));
        store_fe(v + 2 * i, run);
    }
}
// strip-local inclusive prefix products of v (in place) and the strips' totals
__global__ void __launch_bounds__(128) k_prodscan_local(uint4 *__restrict__ v, size_t n, uint4 *__restrict__ totals) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t first = t * PK_SCAN_STRIP;
    if (first >= n) return;
    const size_t end = (first + PK_SCAN_STRIP < n) ? first + PK_SCAN_STRIP : n;
    fe run = load_fe_plain(v + 2 * first);
    for (size_t i = first + 1; i < end; ++i) {
        run = fr_mul(run, load_fe_plain(v + 2 * i));
        store_fe(v + 2 * i, run);
    }
    store_fe(totals + 2 * t, run);
}
// v[i] *= scanned_totals[strip - 1] for every strip but the first
__global__ void __launch_bounds__(128) k_prodscan_fix(uint4 *__restrict__ v, size_t n, const uint4 *__restrict__ scanned_totals) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t first = t * PK_SCAN_STRIP;
    if (t == 0 || first >= n) return;
    const size_t end = (first + PK_SCAN_STRIP < n) ? first + PK_SCAN_STRIP : n;
    const fe cin = load_fe_plain(scanned_totals + 2 * (t - 1));
    for (size_t i = first; i < end; ++i) store_fe(v + 2 * i, fr_mul(cin, load_fe_plain(v + 2 * i)));
}
// Inclusive prefix product of totals[0..n) in place (recursive: strips of 256).  scratch: n / 256 + ... elements.
inline void pk_enqueue_prodscan(uint4 *v, size_t n, uint4 *scratch, pk_stream_t stream) {
    if (n <= 64) {
        PK_LAUNCH(k_prodscan_serial, dim3(1), dim3(32), 0, stream, v, n);
        return;
    }
    const size_t strips = (n + PK_SCAN_STRIP - 1) / PK_SCAN_STRIP;
    const unsigned blocks = (unsigned)((strips + 127) / 128);
    PK_LAUNCH(k_prodscan_local, dim3(blocks), dim3(128), 0, stream, v, n, scratch);
    pk_enqueue_prodscan(scratch, strips, scratch + 2 * strips, stream);
    PK_LAUNCH(k_prodscan_fix, dim3(blocks), dim3(128), 0, stream, v, n, (const uint4 *)scratch);
}
// z polynomials from the scanned sequence: flat index i = c + nc * nth holds 0 (i < nc), 1 (i == nc) or seq[i - nc - 1];
// polynomial c takes it at row bh(nth) (into_bh_order, prover.rs:333-341).
__global__ void __launch_bounds__(128) k_perm_scatter(const uint4 *__restrict__ seq, PermSeq q, size_t rows, uint4 *const *__restrict__ out_polys) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t first = t * PK_SCAN_STRIP;      // rows in hypercube order handled by this thread
    if (first >= rows) return;
    const size_t end = (first + PK_SCAN_STRIP < rows) ? first + PK_SCAN_STRIP : rows;
    fe one;
    one.l[0] = 0x4ffffffbu; one.l[1] = 0xac96341cu; one.l[2] = 0x9f60cd29u; one.l[3] = 0x36fc7695u;
    one.l[4] = 0x7879462eu; one.l[5] = 0x666ea36fu; one.l[6] = 0x9a07df2fu; one.l[7] = 0x0e0a77c1u;
    u32 b = first ? bh_pow_x(first - 1, q.k, q.primitive) : 0u;
    for (size_t nth = first; nth < end; ++nth) {
        if (nth == 1) b = bh_pow_x(0, q.k, q.primitive);
        else if (nth > first && nth > 1) b = bh_next(b, q.k, q.primitive);
        for (u32 c = 0; c < q.nc; ++c) {
            const size_t i = c + (size_t)q.nc * nth;
            fe v;
            if (i < q.nc) v = fe_zero();
            else if (i == q.nc) v = one;
            else v = load_fe_plain(seq + 2 * (i - q.nc - 1));
            store_fe(out_polys[c] + 2 * (size_t)b, v);
        }
    }
}

// util/arithmetic/bh.rs:5-38: the primitive polynomial of every hypercube size.
static const u32 PK_BH_PRIMITIVES[32] = {1u, 3u, 7u, 11u, 19u, 37u, 67u, 131u, 285u, 529u, 1033u, 2053u, 4179u, 8219u, 16427u, 32771u, 65581u, 131081u,
                                         262183u, 524327u, 1048585u, 2097157u, 4194307u, 8388641u, 16777243u, 33554441u, 67108935u, 134217767u,
                                         268435465u, 536870917u, 1073741907u, 2147483657u};
// scratch of pk_enqueue_perm_z in 32-byte elements: products (nc * n) | scanned sequence | strip totals and their recursion
inline size_t pk_perm_z_scratch_elems(size_t num_chunks, size_t n) {
    const size_t m = num_chunks * n - num_chunks - 1;
    const size_t strips = (m + PK_SCAN_STRIP - 1) / PK_SCAN_STRIP;
    return num_chunks * n + m + 2 * strips + 256 + n;  // products, sequence, strip totals (+ their scan), the denominators of one chunk
}
// chunks[k]: the permutation polynomials of chunk k; d_beta_gamma: beta, gamma (two elements, device); d_out_table: device
// array of num_chunks output pointers (2^num_vars elements each).
inline void pk_enqueue_perm_z(const PermArgs *chunks, u32 num_chunks, u32 num_vars, const void *d_beta_gamma, void *scratch, void *d_out_table,
                              pk_stream_t stream) {
    const size_t n = (size_t)1 << num_vars;
    const size_t m = (size_t)num_chunks * n - num_chunks - 1;   // factors entering the flat z sequence (prover.rs:303-320)
    const size_t strips = (m + PK_SCAN_STRIP - 1) / PK_SCAN_STRIP;
    uint4 *d_products = (uint4 *)scratch;
    uint4 *d_seq = d_products + 2 * (size_t)num_chunks * n;
    uint4 *d_totals = d_seq + 2 * m;
    uint4 *d_den = d_totals + 2 * (2 * strips + 256);
    PermSeq seq;
    memset(&seq, 0, sizeof(seq));
    seq.nc = num_chunks; seq.k = num_vars; seq.primitive = PK_BH_PRIMITIVES[num_vars];
    const u32 strip = pk_perm_strip(n);
    const size_t pthreads = (n + strip - 1) / strip;
    for (u32 k = 0; k < num_chunks; ++k) {
        seq.product[k] = d_products + 2 * (size_t)k * n;
        PK_LAUNCH(k_perm_products, dim3((unsigned)((pthreads + 127) / 128)), dim3(128), 0, stream, chunks[k], (const uint4 *)d_beta_gamma, n, strip, d_den, d_products + 2 * (size_t)k * n);
    }
    if (m) {
        PK_LAUNCH(k_perm_gather_scan, dim3((unsigned)((strips + 127) / 128)), dim3(128), 0, stream, seq, m, d_seq, d_totals);
        if (strips > 1) {
            pk_enqueue_prodscan(d_totals, strips, d_totals + 2 * strips, stream);
            PK_LAUNCH(k_prodscan_fix, dim3((unsigned)((strips + 127) / 128)), dim3(128), 0, stream, d_seq, m, (const uint4 *)d_totals);
        }
    }
    const size_t sthreads = (n + PK_SCAN_STRIP - 1) / PK_SCAN_STRIP;
    PK_LAUNCH(k_perm_scatter, dim3((unsigned)((sthreads + 127) / 128)), dim3(128), 0, stream, (const uint4 *)d_seq, seq, n, (uint4 *const *)d_out_table);
}

// ------------------------------------------------- affine tables for the sum-check expression compiler
// The reference's sum check keeps identity / Lagrange polynomials, constants and rotated queries implicit in its
// expression evaluator (piop/sum_check/classic.rs:40-75, 104-126; classic/eval.rs).  Every sub-expression of degree <= 1
//     constant + id_coeff * identity + sum_i coeff_i * poly_i(rotated by r_i)         (+ a few single-row terms: Lagrange)
// is a multilinear polynomial itself and fixing a variable commutes with the sum, so the compiler (plonkish_b200/
// expression.py) materialises it as one table: the round values stay the same field elements.  The row of a rotated
// query is BooleanHypercube::rotate (util/arithmetic/bh.rs:104-121), what classic.rs:105-125 gathers through
// rotation_map.  HBM bound: 32 B written + 32 B per source read.
#define PK_AFFINE_MAX 8
static const u32 PK_BH_X_INVS[32] = {0u, 1u, 3u, 5u, 9u, 18u, 33u, 65u, 142u, 264u, 516u, 1026u, 2089u, 4109u, 8213u, 16385u, 32790u, 65540u,
                                     131091u, 262163u, 524292u, 1048578u, 2097153u, 4194320u, 8388621u, 16777220u, 33554467u, 67108883u,
                                     134217732u, 268435458u, 536870953u, 1073741828u};
struct AffineArgs {
    const uint4 *poly[PK_AFFINE_MAX];
    fe coeff[PK_AFFINE_MAX];
    int rotation[PK_AFFINE_MAX];
    fe constant, id_coeff;  // Montgomery Fr
    u32 count, has_constant, has_id, has_coeff_mask;  // bit i of has_coeff_mask: coeff[i] != 1
    u32 num_vars, primitive, x_inv;
};
PK_HD u32 bh_rotate(u32 b, int rotation, u32 k, u32 primitive, u32 x_inv) {  // bh.rs:104-121, 141-153
    for (; rotation > 0; --rotation) b = bh_next(b, k, primitive);
    for (; rotation < 0; ++rotation) b = (b >> 1) ^ ((b & 1u) * x_inv);
    return b;
}
__global__ void __launch_bounds__(256) k_fr_affine(AffineArgs a, size_t n, uint4 *out) {  // `out` may be a source too (unrotated)
    const size_t first = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    // the identity term walks with the grid stride: id_coeff * j once, then one addition per row
    fe id_term = fe_zero(), id_step = fe_zero();
    if (a.has_id) {
        id_term = fr_mul(a.id_coeff, fr_from_u64(first));
        id_step = fr_mul(a.id_coeff, fr_from_u64(stride));
    }
    for (size_t j = first; j < n; j += stride) {
        fe acc = a.has_constant ? a.constant : fe_zero();
        if (a.has_id) {
            acc = fr_add(acc, id_term);
            id_term = fr_add(id_term, id_step);
        }
        for (u32 i = 0; i < a.count; ++i) {
            const size_t row = a.rotation[i] ? (size_t)bh_rotate((u32)j, a.rotation[i], a.num_vars, a.primitive, a.x_inv) : j;
            const fe v = load_fe_plain(a.poly[i] + 2 * row);
            acc = fr_add(acc, ((a.has_coeff_mask >> i) & 1u) ? fr_mul(a.coeff[i], v) : v);
        }
        store_fe(out + 2 * j, acc);
    }
}
// out[rows[i]] += values[i] (Lagrange terms, instance polynomials: a handful of rows; rows are distinct or serialised by
// the single thread).
__global__ void k_fr_sparse_add(uint4 *__restrict__ out, const unsigned long long *__restrict__ rows, const uint4 *__restrict__ values, u32 count) {
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (u32 i = 0; i < count; ++i) store_fe(out + 2 * rows[i], fr_add(load_fe_plain(out + 2 * rows[i]), load_fe_plain(values + 2 * (size_t)i)));
}

// ------------------------------------------------------------- fixed-base MSM
// Signed 16-bit windows: 16 windows cover 254 bits plus the carry, the table holds
// d * 2^(16w) * base for d = 1..2^15 (16 x 32768 x 64 B = 32 MiB, L2 resident); the
// reference's unsigned table (msm.rs:16-31) is the same set of multiples, twice as many.
#define PK_FIXED_C 16
#define PK_FIXED_W 16
#define PK_FIXED_ROW (1u << (PK_FIXED_C - 1))

// offsets[w] = 2^(16w) * base, affine (the `offset` of msm.rs:22).  One thread.
__global__ void k_fixed_offsets(const affine *__restrict__ base, affine *__restrict__ offsets) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    xyzz cur = xyzz_from_affine(*base);
    for (u32 w = 0; w < PK_FIXED_W; ++w) {
        offsets[w] = xyzz_to_affine(cur);
        for (u32 b = 0; b < PK_FIXED_C; ++b) cur = xyzz_double(cur);
    }
}
// rows[w][d-1] = d * offsets[w] (projective; normalised by k_table_normalize afterwards).
__global__ void __launch_bounds__(128) k_fixed_table(const affine *__restrict__ offsets, xyzz *__restrict__ rows) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= PK_FIXED_W * PK_FIXED_ROW) return;
    const u32 w = t / PK_FIXED_ROW, d = t % PK_FIXED_ROW + 1;
    const affine off = offsets[w];
    xyzz acc = xyzz_identity();
    for (int bit = 31 - __clz(d); bit >= 0; --bit) {
        acc = xyzz_double(acc);
        if ((d >> bit) & 1) xyzz_madd(acc, off.x, off.y);
    }
    store_xyzz(rows + t, acc);
}
// out[i] = scalars[i] * base (windowed_scalar_mul, msm.rs:50-65, with signed digits).
__global__ void __launch_bounds__(128, 4) k_fixed_base_mul(const uint4 *__restrict__ scalars, u32 n, const affine *__restrict__ table,
                                                           xyzz *__restrict__ out) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fe v = fr_to_canonical(load_fe(scalars + 2 * (size_t)i));  // to_repr, msm.rs:56
    xyzz acc = xyzz_identity();
    u32 carry = 0;
    for (u32 w = 0; w < PK_FIXED_W; ++w) {
        const u32 raw = ((v.l[w >> 1] >> ((w & 1) * 16)) & 0xffffu) + carry;
        carry = raw > PK_FIXED_ROW ? 1u : 0u;
        const u32 mag = carry ? (1u << PK_FIXED_C) - raw : raw;
        if (mag == 0) continue;
        const uint4 *tp = reinterpret_cast<const uint4 *>(table + (size_t)w * PK_FIXED_ROW + (mag - 1));
        const fe x = load_fe(tp);
        fe y = load_fe(tp + 2);
        if (carry) y = fq_neg(y);
        xyzz_madd<MulInline>(acc, x, y);
    }
    store_xyzz(out + i, acc);
}

// Table for one base: offsets, multiples, normalise.  tmp holds 16 x 32768 xyzz (64 MiB).
inline void pk_enqueue_fixed_table(const void *d_base, affine *d_offsets, xyzz *tmp, affine *table, pk_stream_t stream) {
    const u32 total = PK_FIXED_W * PK_FIXED_ROW;
    PK_LAUNCH(k_fixed_offsets, dim3(1), dim3(32), 0, stream, (const affine *)d_base, d_offsets);
    PK_LAUNCH(k_fixed_table, dim3((total + 127) / 128), dim3(128), 0, stream, (const affine *)d_offsets, tmp);
    PK_LAUNCH(k_table_normalize, dim3(((total + 7) / 8 + 127) / 128), dim3(128), 0, stream, (const xyzz *)tmp, total, table);
}
// out_affine[i] = scalars[i] * base for n <= 2^22 scalars per call; tmp holds n xyzz.
inline void pk_enqueue_fixed_base(const void *d_scalars, u32 n, const affine *table, xyzz *tmp, affine *out_affine, pk_stream_t stream) {
    PK_LAUNCH(k_fixed_base_mul, dim3((n + 127) / 128), dim3(128), 0, stream, (const uint4 *)d_scalars, n, table, tmp);
    PK_LAUNCH(k_table_normalize, dim3(((n + 7) / 8 + 127) / 128), dim3(128), 0, stream, (const xyzz *)tmp, n, out_affine);
}

// ------------------------------------------------------------------ Zeromorph
// Zeromorph<UnivariateKzg>::open (pcs/multilinear/zeromorph.rs:126-186) builds two univariate polynomials of 2^n
// coefficients out of the quotients of `quotients` (pcs/multilinear.rs:72-107, q_i of 2^i values packed at element
// offset 2^i of `q`, as pk_enqueue_quotients leaves them).  Both are one pass over 2^n elements with 2^n products in
// total (element m of either sum has as many terms as there are quotients long enough to reach it): one product per
// 64 B moved, which on B200 is multiplier bound rather than HBM bound (profiles/r02_pcs_field_kernels_summary.txt).
#define PK_ZM_MAX_VARS 28
struct ZmWeights {
    fe w[PK_ZM_MAX_VARS];
};
// q_hat (zeromorph.rs:157-167): q_hat[2^n - 2^i + j] += y^i * q_i[j] for j < 2^i, every quotient aligned to the top.
// With d = 2^n - m:  q_hat[m] = sum_{i < n, 2^i >= d} w[i] * q[2^(i+1) - d],  w[i] = y^i.
__global__ void __launch_bounds__(256) k_zm_q_hat(const uint4 *__restrict__ q, ZmWeights a, u32 num_vars, uint4 *__restrict__ out) {
    const size_t n = (size_t)1 << num_vars;
    for (size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x; m < n; m += (size_t)gridDim.x * blockDim.x) {
        const size_t d = n - m;                                   // 1 <= d <= 2^28
        const u32 i0 = d > 1 ? 32u - (u32)__clz((u32)(d - 1)) : 0u;  // smallest i with 2^i >= d
        fe acc = fe_zero();
        for (u32 i = i0; i < num_vars; ++i) acc = fr_add(acc, fr_mul(a.w[i], load_fe_plain(q + 2 * (((size_t)2 << i) - d))));
        store_fe(out + 2 * m, acc);
    }
}
// f (zeromorph.rs:175-180): f = z * poly + q_hat; f[0] += c0 (= eval_scalar * eval); f[j] += s_i * q_i[j] for j < 2^i,
// every quotient aligned to the bottom:  f[m] = z poly[m] + q_hat[m] + sum_{i < n, 2^i > m} w[i] * q[2^i + m],  w[i] = s_i.
__global__ void __launch_bounds__(256) k_zm_f(const uint4 *__restrict__ poly, const uint4 *__restrict__ q_hat, const uint4 *__restrict__ q,
                                              ZmWeights a, fe z, fe c0, u32 num_vars, uint4 *__restrict__ out) {
    const size_t n = (size_t)1 << num_vars;
    for (size_t m = blockIdx.x * (size_t)blockDim.x + threadIdx.x; m < n; m += (size_t)gridDim.x * blockDim.x) {
        fe acc = fr_add(fr_mul(z, load_fe(poly + 2 * m)), load_fe_plain(q_hat + 2 * m));
        if (m == 0) acc = fr_add(acc, c0);
        const u32 i0 = m ? 32u - (u32)__clz((u32)m) : 0u;            // smallest i with 2^i > m (m < 2^28)
        for (u32 i = i0; i < num_vars; ++i) acc = fr_add(acc, fr_mul(a.w[i], load_fe_plain(q + 2 * (((size_t)1 << i) + m))));
        store_fe(out + 2 * m, acc);
    }
}
inline u32 pk_zm_blocks(u32 num_vars, u32 sm_count) {
    const size_t n = (size_t)1 << num_vars;
    size_t blocks = (n + 255) / 256;
    const size_t cap = (size_t)sm_count * 8;
    return (u32)(blocks > cap ? cap : blocks);
}
inline void pk_enqueue_zm_q_hat(const void *q, const ZmWeights &w, u32 num_vars, void *out, u32 sm_count, pk_stream_t stream) {
    PK_LAUNCH(k_zm_q_hat, dim3(pk_zm_blocks(num_vars, sm_count)), dim3(256), 0, stream, (const uint4 *)q, w, num_vars, (uint4 *)out);
}
inline void pk_enqueue_zm_f(const void *poly, const void *q_hat, const void *q, const ZmWeights &w, const fe &z, const fe &c0, u32 num_vars,
                            void *out, u32 sm_count, pk_stream_t stream) {
    PK_LAUNCH(k_zm_f, dim3(pk_zm_blocks(num_vars, sm_count)), dim3(256), 0, stream, (const uint4 *)poly, (const uint4 *)q_hat, (const uint4 *)q, w, z,
              c0, num_vars, (uint4 *)out);
}

}  // namespace pk
