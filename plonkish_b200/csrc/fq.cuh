// BN254 field arithmetic for sm_100a: 254-bit Montgomery values on 8 x 32-bit
// limbs, products built from mad.lo.cc / madc.hi.cc carry chains (ptxas fuses each
// lo/hi pair into one IMAD.WIDE.U32 with carry-in/out), even/odd column
// accumulators so no carry ever ripples between the two halves of a product.
//
// Replaces, on the device, the halo2_curves 0.3.3 [ext] Fq/Fr arithmetic that
// /root/reference/plonkish_backend/src/util/arithmetic/msm.rs:133-148,153,163
// calls into (64-bit-limb mac/adc on the CPU).
//
// The same file compiles under g++ with tests/emul/cuda_emul.h, where every
// carry-chain block below has a portable restatement; that is test plumbing —
// the product only ever runs the PTX.
#pragma once
#include <stdint.h>

#ifndef PLONKISH_EMUL
#define PK_HD __device__ __forceinline__
#else
#define PK_HD inline
#endif

namespace pk {

typedef uint32_t u32;
typedef uint64_t u64;

struct fe {  // one field element, little-endian 32-bit limbs, Montgomery form (R = 2^256)
    u32 l[8];
};

// ---------------------------------------------------------------- constants
// Fq modulus p, -p^-1 mod 2^32, R mod p (the Montgomery "one").
#define PK_P0 0xd87cfd47u
#define PK_P1 0x3c208c16u
#define PK_P2 0x6871ca8du
#define PK_P3 0x97816a91u
#define PK_P4 0x8181585du
#define PK_P5 0xb85045b6u
#define PK_P6 0xe131a029u
#define PK_P7 0x30644e72u
#define PK_PINV 0xe4866389u
// Fr modulus r and -r^-1 mod 2^32.
#define PK_R0 0xf0000001u
#define PK_R1 0x43e1f593u
#define PK_R2 0x79b97091u
#define PK_R3 0x2833e848u
#define PK_R4 0x8181585du
#define PK_R5 0xb85045b6u
#define PK_R6 0xe131a029u
#define PK_R7 0x30644e72u
#define PK_RINV 0xefffffffu

PK_HD fe fq_one() {
    fe r = {{0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}};
    return r;
}
PK_HD fe fe_zero() {
    fe r = {{0, 0, 0, 0, 0, 0, 0, 0}};
    return r;
}
PK_HD bool fe_is_zero(const fe &a) {
    return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0;
}
PK_HD bool fe_eq(const fe &a, const fe &b) {
    u32 d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.l[i] ^ b.l[i];
    return d == 0;
}

// ------------------------------------------------------- carry-chain blocks
// Each block is ONE asm statement, so the condition-code register never has to
// survive between statements.  MOD selects the modulus for the immediate forms.

#ifndef PLONKISH_EMUL

// acc[0..7] += {x0,x2,x4,x6} * y laid out as 4 (lo,hi) pairs; returns carry-out.
PK_HD u32 cmad8(u32 *acc, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) {
    u32 c;
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "=r"(c)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    return c;
}

// The same with the carry-out added to `top` inside the chain (one IADD3.X; returning the carry as a value and adding
// it afterwards costs a predicated add pair and a move: the accumulate loop is bound by instruction issue, where every
// ALU instruction next to the multiplies costs about a cycle — plonkish_cuda_bench_issue_mix).
PK_HD void cmad8_top(u32 *acc, u32 &top, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) {
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(top)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
}
// ... and for a chain whose top pair has headroom: no carry-out at all.
PK_HD void cmad8_nc(u32 *acc, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) {
    asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32    %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
}

// e0 += carry_word; then o[k] = o[k+2] + {x1,x3,x5,x7} * y pairs (o[8] = o[9] = 0), the
// carry of the first add entering the chain.  The top pair cannot overflow because
// x7 < 2^31 (operands < 2p < 2^255).
PK_HD void shift_mad8(u32 &e0, u32 carry_word, u32 *o, u32 x1, u32 x3, u32 x5, u32 x7, u32 y) {
    asm("add.cc.u32     %0, %0, %9;\n\t"
        "madc.lo.cc.u32 %1, %10, %14, %3;\n\t"
        "madc.hi.cc.u32 %2, %10, %14, %4;\n\t"
        "madc.lo.cc.u32 %3, %11, %14, %5;\n\t"
        "madc.hi.cc.u32 %4, %11, %14, %6;\n\t"
        "madc.lo.cc.u32 %5, %12, %14, %7;\n\t"
        "madc.hi.cc.u32 %6, %12, %14, %8;\n\t"
        "madc.lo.cc.u32 %7, %13, %14, 0;\n\t"
        "madc.hi.u32    %8, %13, %14, 0;"
        : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
        : "r"(carry_word), "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
}

// out = a + b (8 limbs), returns carry-out.
PK_HD u32 add8(u32 *out, const u32 *a, const u32 *b) {
    u32 c;
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=r"(out[0]), "=r"(out[1]), "=r"(out[2]), "=r"(out[3]), "=r"(out[4]), "=r"(out[5]), "=r"(out[6]),
          "=r"(out[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
}

// out = a - b (8 limbs), returns borrow (0 or 1).
PK_HD u32 sub8(u32 *out, const u32 *a, const u32 *b) {
    u32 c;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=r"(out[0]), "=r"(out[1]), "=r"(out[2]), "=r"(out[3]), "=r"(out[4]), "=r"(out[5]), "=r"(out[6]),
          "=r"(out[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c & 1u;
}

#else  // PLONKISH_EMUL: portable restatements of the blocks above (tests only)

inline u32 cmad8(u32 *acc, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) {
    const u32 x[4] = {x0, x2, x4, x6};
    u64 carry = 0;
    for (int k = 0; k < 4; ++k) {
        u64 prod = (u64)x[k] * y;
        u64 lo = (u64)acc[2 * k] + (u32)prod + carry;
        acc[2 * k] = (u32)lo;
        u64 hi = (u64)acc[2 * k + 1] + (u32)(prod >> 32) + (lo >> 32);
        acc[2 * k + 1] = (u32)hi;
        carry = hi >> 32;
    }
    return (u32)carry;
}
inline void cmad8_top(u32 *acc, u32 &top, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) { top += cmad8(acc, x0, x2, x4, x6, y); }
inline void cmad8_nc(u32 *acc, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) { (void)cmad8(acc, x0, x2, x4, x6, y); }
inline void shift_mad8(u32 &e0, u32 carry_word, u32 *o, u32 x1, u32 x3, u32 x5, u32 x7, u32 y) {
    const u32 x[4] = {x1, x3, x5, x7};
    u64 t = (u64)e0 + carry_word;
    e0 = (u32)t;
    u64 carry = t >> 32;
    for (int k = 0; k < 4; ++k) {
        u64 prod = (u64)x[k] * y;
        u32 src_lo = (2 * k + 2 < 8) ? o[2 * k + 2] : 0, src_hi = (2 * k + 3 < 8) ? o[2 * k + 3] : 0;
        u64 lo = (u64)src_lo + (u32)prod + carry;
        o[2 * k] = (u32)lo;
        u64 hi = (u64)src_hi + (u32)(prod >> 32) + (lo >> 32);
        o[2 * k + 1] = (u32)hi;
        carry = hi >> 32;
    }
}
inline u32 add8(u32 *out, const u32 *a, const u32 *b) {
    u64 carry = 0;
    for (int i = 0; i < 8; ++i) {
        u64 t = (u64)a[i] + b[i] + carry;
        out[i] = (u32)t;
        carry = t >> 32;
    }
    return (u32)carry;
}
inline u32 sub8(u32 *out, const u32 *a, const u32 *b) {
    u64 borrow = 0;
    for (int i = 0; i < 8; ++i) {
        u64 t = (u64)a[i] - b[i] - borrow;
        out[i] = (u32)t;
        borrow = (t >> 32) & 1;
    }
    return (u32)borrow;
}
#endif

// ------------------------------------------------------------ modulus tables
struct FqMod {
    static PK_HD u32 inv() { return PK_PINV; }
    static PK_HD void limbs(u32 *m) {
        m[0] = PK_P0; m[1] = PK_P1; m[2] = PK_P2; m[3] = PK_P3;
        m[4] = PK_P4; m[5] = PK_P5; m[6] = PK_P6; m[7] = PK_P7;
    }
};
struct FrMod {
    static PK_HD u32 inv() { return PK_RINV; }
    static PK_HD void limbs(u32 *m) {
        m[0] = PK_R0; m[1] = PK_R1; m[2] = PK_R2; m[3] = PK_R3;
        m[4] = PK_R4; m[5] = PK_R5; m[6] = PK_R6; m[7] = PK_R7;
    }
};

// r = (r >= m) ? r - m : r, for r < 2m.
template <class MOD>
PK_HD void final_sub(u32 *r) {
    u32 m[8], t[8];
    MOD::limbs(m);
    u32 borrow = sub8(t, r, m);
    if (!borrow) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = t[i];
    }
}

// Montgomery product a*b/2^256 mod m, a,b < m, result < m.
//
// Word-serial reduction, one row per limb of b.  The running value T is kept
// as two accumulators: `ev` holds the 64-bit partial products that start at even
// word positions, `od` those that start at odd positions (its word k sits at
// true word k+1).  A row adds a*b_i, then m*q with q = T[0]*(-m^-1), which clears
// T[0]; dividing by 2^32 turns the odd accumulator into the even one as is, and
// the old even one (minus its dead word 0) into the odd one — the single
// left-over word ev[1] is folded in by shift_mad8's first add.  136 multiply
// instructions per product: 64 (a*b) + 64 (m*q) + 8 (q).
//
// LAZY = true is the redundant-range form the accumulate loop uses: operands and result in [0, 2m) and no final
// subtraction (a*b/R + m < 4m^2/R + m < 2m because 4m < R for both moduli; the running value stays below a + m < 3m).
template <class MOD, bool LAZY = false>
PK_HD fe mont_mul(const fe &a, const fe &b) {
    u32 m[8];
    MOD::limbs(m);
    u32 ev[8], od[8];

    // row 0
    {
        const u32 y = b.l[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u64 pe = (u64)a.l[2 * k] * y, po = (u64)a.l[2 * k + 1] * y;
            ev[2 * k] = (u32)pe; ev[2 * k + 1] = (u32)(pe >> 32);
            od[2 * k] = (u32)po; od[2 * k + 1] = (u32)(po >> 32);
        }
        const u32 q = ev[0] * MOD::inv();
        cmad8_nc(od, m[1], m[3], m[5], m[7], q);               // top pair has headroom: no carry-out
        cmad8_top(ev, od[7], m[0], m[2], m[4], m[6], q);      // carry sits at true word 8 = od word 7
    }
    // rows 1..7, roles of ev/od alternate
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        u32 *E = (i & 1) ? od : ev;   // aligned at even positions after the shift
        u32 *O = (i & 1) ? ev : od;   // old even accumulator: word 0 is dead, word 1 folds into E[0]
        const u32 y = b.l[i];
        shift_mad8(E[0], O[1], O, a.l[1], a.l[3], a.l[5], a.l[7], y);
        cmad8_top(E, O[7], a.l[0], a.l[2], a.l[4], a.l[6], y);
        const u32 q = E[0] * MOD::inv();
        cmad8_nc(O, m[1], m[3], m[5], m[7], q);
        cmad8_top(E, O[7], m[0], m[2], m[4], m[6], q);
    }
    // Row 7 ran with E = od, O = ev and left od[0] == 0; the last shift gives
    // T/2^256 = ev + (od >> 32).
    fe r;
    {
        u32 sh[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) sh[k] = od[k + 1];
        sh[7] = 0;
        add8(r.l, ev, sh);
    }
    if (!LAZY) final_sub<MOD>(r.l);
    return r;
}

// (a*b + c*d) / 2^256 mod m with ONE reduction: every row adds both partial products before
// the m*q step, 200 multiply instructions instead of 272 for two products and an addition.
// Bounds: T_{i+1} < T_i / 2^32 + (a + c + m)(1 - 2^-32) keeps T < 3m - 3 < 2^256 for a, c < m, so
// the accumulators never carry out of their top pairs (3 * m7 * 2^32 < 2^64 for both moduli) and
// two conditional subtractions finish.
// LAZY: operands in [0, 2m), running value below a + c + m < 5m < 2^256, result below 8m^2/R + m < 2.6m: one conditional
// subtraction brings it under 2m.
template <class MOD, bool LAZY = false>
PK_HD fe mont_mul_sum(const fe &a, const fe &b, const fe &c, const fe &d) {
    u32 m[8];
    MOD::limbs(m);
    u32 ev[8], od[8];
    {
        const u32 y = b.l[0], z = d.l[0];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u64 pe = (u64)a.l[2 * k] * y, po = (u64)a.l[2 * k + 1] * y;
            ev[2 * k] = (u32)pe; ev[2 * k + 1] = (u32)(pe >> 32);
            od[2 * k] = (u32)po; od[2 * k + 1] = (u32)(po >> 32);
        }
        cmad8_nc(od, c.l[1], c.l[3], c.l[5], c.l[7], z);
        cmad8_top(ev, od[7], c.l[0], c.l[2], c.l[4], c.l[6], z);
        const u32 q = ev[0] * MOD::inv();
        cmad8_nc(od, m[1], m[3], m[5], m[7], q);
        cmad8_top(ev, od[7], m[0], m[2], m[4], m[6], q);
    }
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        u32 *E = (i & 1) ? od : ev;
        u32 *O = (i & 1) ? ev : od;
        const u32 y = b.l[i], z = d.l[i];
        shift_mad8(E[0], O[1], O, a.l[1], a.l[3], a.l[5], a.l[7], y);
        cmad8_top(E, O[7], a.l[0], a.l[2], a.l[4], a.l[6], y);
        cmad8_nc(O, c.l[1], c.l[3], c.l[5], c.l[7], z);
        cmad8_top(E, O[7], c.l[0], c.l[2], c.l[4], c.l[6], z);
        const u32 q = E[0] * MOD::inv();
        cmad8_nc(O, m[1], m[3], m[5], m[7], q);
        cmad8_top(E, O[7], m[0], m[2], m[4], m[6], q);
    }
    fe r;
    {
        u32 sh[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) sh[k] = od[k + 1];
        sh[7] = 0;
        add8(r.l, ev, sh);
    }
    final_sub<MOD>(r.l);
    if (!LAZY) final_sub<MOD>(r.l);
    return r;
}


// ---- variants of the two row blocks that skip the partial products of the first S limb pairs (squaring: the
// symmetric products of row i are taken in the earlier rows).  Skipped pairs of the shifted accumulator still move
// down two words and carry on, with add-with-carry on the ALU pipe instead of a multiply.
#ifndef PLONKISH_EMUL
template <int S>
PK_HD u32 cmad8_from(u32 *acc, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) {
    u32 c = 0;
    if constexpr (S == 0) {
        asm(
            "mad.lo.cc.u32  %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, 0, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    }
    else if constexpr (S == 1) {
        asm(
            "mad.lo.cc.u32  %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, 0, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    }
    else if constexpr (S == 2) {
        asm(
            "mad.lo.cc.u32  %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, 0, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    }
    else if constexpr (S == 3) {
        asm(
            "mad.lo.cc.u32  %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, 0, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(c)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    }
    return c;  // S == 4: nothing to add
}
template <int S>
PK_HD void cmad8_from_top(u32 *acc, u32 &top, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) {
    if constexpr (S == 0) {
        asm(
            "mad.lo.cc.u32  %0, %9, %13, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
            "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    }
    else if constexpr (S == 1) {
        asm(
            "mad.lo.cc.u32  %2, %10, %13, %2;\n\t"
            "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
            "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    }
    else if constexpr (S == 2) {
        asm(
            "mad.lo.cc.u32  %4, %11, %13, %4;\n\t"
            "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
            "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    }
    else if constexpr (S == 3) {
        asm(
            "mad.lo.cc.u32  %6, %12, %13, %6;\n\t"
            "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
            "addc.u32       %8, %8, 0;"
            : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
            : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y));
    }
}
template <int S>
PK_HD void shift_mad8_from(u32 &e0, u32 carry_word, u32 *o, u32 x1, u32 x3, u32 x5, u32 x7, u32 y) {
    if constexpr (S == 0) {
        asm(
            "add.cc.u32     %0, %0, %9;\n\t"
            "madc.lo.cc.u32 %1, %10, %14, %3;\n\t"
            "madc.hi.cc.u32 %2, %10, %14, %4;\n\t"
            "madc.lo.cc.u32 %3, %11, %14, %5;\n\t"
            "madc.hi.cc.u32 %4, %11, %14, %6;\n\t"
            "madc.lo.cc.u32 %5, %12, %14, %7;\n\t"
            "madc.hi.cc.u32 %6, %12, %14, %8;\n\t"
            "madc.lo.cc.u32 %7, %13, %14, 0;\n\t"
            "madc.hi.u32    %8, %13, %14, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(carry_word), "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
    }
    else if constexpr (S == 1) {
        asm(
            "add.cc.u32     %0, %0, %9;\n\t"
            "addc.cc.u32    %1, %3, 0;\n\t"
            "addc.cc.u32    %2, %4, 0;\n\t"
            "madc.lo.cc.u32 %3, %11, %14, %5;\n\t"
            "madc.hi.cc.u32 %4, %11, %14, %6;\n\t"
            "madc.lo.cc.u32 %5, %12, %14, %7;\n\t"
            "madc.hi.cc.u32 %6, %12, %14, %8;\n\t"
            "madc.lo.cc.u32 %7, %13, %14, 0;\n\t"
            "madc.hi.u32    %8, %13, %14, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(carry_word), "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
    }
    else if constexpr (S == 2) {
        asm(
            "add.cc.u32     %0, %0, %9;\n\t"
            "addc.cc.u32    %1, %3, 0;\n\t"
            "addc.cc.u32    %2, %4, 0;\n\t"
            "addc.cc.u32    %3, %5, 0;\n\t"
            "addc.cc.u32    %4, %6, 0;\n\t"
            "madc.lo.cc.u32 %5, %12, %14, %7;\n\t"
            "madc.hi.cc.u32 %6, %12, %14, %8;\n\t"
            "madc.lo.cc.u32 %7, %13, %14, 0;\n\t"
            "madc.hi.u32    %8, %13, %14, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(carry_word), "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
    }
    else if constexpr (S == 3) {
        asm(
            "add.cc.u32     %0, %0, %9;\n\t"
            "addc.cc.u32    %1, %3, 0;\n\t"
            "addc.cc.u32    %2, %4, 0;\n\t"
            "addc.cc.u32    %3, %5, 0;\n\t"
            "addc.cc.u32    %4, %6, 0;\n\t"
            "addc.cc.u32    %5, %7, 0;\n\t"
            "addc.cc.u32    %6, %8, 0;\n\t"
            "madc.lo.cc.u32 %7, %13, %14, 0;\n\t"
            "madc.hi.u32    %8, %13, %14, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(carry_word), "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
    }
    else if constexpr (S == 4) {
        asm(
            "add.cc.u32     %0, %0, %9;\n\t"
            "addc.cc.u32    %1, %3, 0;\n\t"
            "addc.cc.u32    %2, %4, 0;\n\t"
            "addc.cc.u32    %3, %5, 0;\n\t"
            "addc.cc.u32    %4, %6, 0;\n\t"
            "addc.cc.u32    %5, %7, 0;\n\t"
            "addc.cc.u32    %6, %8, 0;\n\t"
            "addc.cc.u32    %7, 0, 0;\n\t"
            "addc.u32       %8, 0, 0;"
            : "+r"(e0), "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7])
            : "r"(carry_word), "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(y));
    }
}
#else
template <int S>
inline u32 cmad8_from(u32 *acc, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) {
    return cmad8(acc, S > 0 ? 0u : x0, S > 1 ? 0u : x2, S > 2 ? 0u : x4, S > 3 ? 0u : x6, y);
}
template <int S>
inline void cmad8_from_top(u32 *acc, u32 &top, u32 x0, u32 x2, u32 x4, u32 x6, u32 y) { top += cmad8_from<S>(acc, x0, x2, x4, x6, y); }
template <int S>
inline void shift_mad8_from(u32 &e0, u32 carry_word, u32 *o, u32 x1, u32 x3, u32 x5, u32 x7, u32 y) {
    shift_mad8(e0, carry_word, o, S > 0 ? 0u : x1, S > 1 ? 0u : x3, S > 2 ? 0u : x5, S > 3 ? 0u : x7, y);
}
#endif

// One row i >= 1 of the squaring: T += a_i * (a_i at word i, the doubled tail 2 * (a >> 32(i+1)) above it), then one
// reduction step.  x[] holds the row's multiplicand limbs (entries below i unused).
template <class MOD, int I>
PK_HD void sqr_row(u32 *ev, u32 *od, const u32 *x, u32 y, const u32 *m) {
    u32 *E = (I & 1) ? od : ev;
    u32 *O = (I & 1) ? ev : od;
    shift_mad8_from<I / 2>(E[0], O[1], O, x[1], x[3], x[5], x[7], y);       // odd limbs below I skipped
    cmad8_from_top<(I + 1) / 2>(E, O[7], x[0], x[2], x[4], x[6], y);          // even limbs below I skipped
    const u32 q = E[0] * MOD::inv();
    cmad8_nc(O, m[1], m[3], m[5], m[7], q);
    cmad8_top(E, O[7], m[0], m[2], m[4], m[6], q);
}

// a*a / 2^256 mod m, a < m: 36 + 64 wide multiplies instead of 128.  Row i multiplies a_i by a_i and by the limbs of
// the doubled tail 2 * (a >> 32(i+1)) only; the running value stays below 3m < 2^256 (a row adds at most 2^32 * 2a),
// the final value below 2m.
// LAZY as in mont_mul: a in [0, 2m) (2a < 2^256 still fits the eight limbs of d), result below 2m unreduced.
template <class MOD, bool LAZY = false>
PK_HD fe mont_sqr(const fe &a) {
    u32 m[8], d[8];
    MOD::limbs(m);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = (a.l[i] << 1) | (i ? (a.l[i - 1] >> 31) : 0u);  // limbs of 2a (a < 2^254)
    u32 ev[8], od[8];
    {   // row 0: every product
        const u32 y = a.l[0];
        const u32 x[8] = {a.l[0], a.l[1] << 1, d[2], d[3], d[4], d[5], d[6], d[7]};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u64 pe = (u64)x[2 * k] * y, po = (u64)x[2 * k + 1] * y;
            ev[2 * k] = (u32)pe; ev[2 * k + 1] = (u32)(pe >> 32);
            od[2 * k] = (u32)po; od[2 * k + 1] = (u32)(po >> 32);
        }
        const u32 q = ev[0] * MOD::inv();
        cmad8_nc(od, m[1], m[3], m[5], m[7], q);
        cmad8_top(ev, od[7], m[0], m[2], m[4], m[6], q);
    }
#define PK_SQR_ROW(I)                                                                                         \
    {                                                                                                         \
        u32 x[8];                                                                                             \
        _Pragma("unroll") for (int j = 0; j < 8; ++j) x[j] = (j == I) ? a.l[I] : (j == I + 1) ? (a.l[j] << 1) : d[j]; \
        sqr_row<MOD, I>(ev, od, x, a.l[I], m);                                                                \
    }
    PK_SQR_ROW(1) PK_SQR_ROW(2) PK_SQR_ROW(3) PK_SQR_ROW(4) PK_SQR_ROW(5) PK_SQR_ROW(6) PK_SQR_ROW(7)
#undef PK_SQR_ROW
    fe r;
    {
        u32 sh[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) sh[k] = od[k + 1];
        sh[7] = 0;
        add8(r.l, ev, sh);
    }
    if (!LAZY) final_sub<MOD>(r.l);
    return r;
}

PK_HD fe fq_mul(const fe &a, const fe &b) { return mont_mul<FqMod>(a, b); }
// a*b + c*d (Montgomery), one reduction
PK_HD fe fq_mul_sum(const fe &a, const fe &b, const fe &c, const fe &d) { return mont_mul_sum<FqMod>(a, b, c, d); }
PK_HD fe fq_sqr(const fe &a) { return mont_sqr<FqMod>(a); }

template <class MOD>
PK_HD fe mod_add(const fe &a, const fe &b) {
    fe r;
    add8(r.l, a.l, b.l);  // a + b < 2^255: no carry-out
    final_sub<MOD>(r.l);
    return r;
}
template <class MOD>
PK_HD fe mod_sub(const fe &a, const fe &b) {
    fe r;
    u32 borrow = sub8(r.l, a.l, b.l);
    if (borrow) {
        u32 m[8];
        MOD::limbs(m);
        add8(r.l, r.l, m);
    }
    return r;
}
PK_HD fe fq_add(const fe &a, const fe &b) { return mod_add<FqMod>(a, b); }
PK_HD fe fq_sub(const fe &a, const fe &b) { return mod_sub<FqMod>(a, b); }
PK_HD fe fq_dbl(const fe &a) { return mod_add<FqMod>(a, a); }
PK_HD fe fq_neg(const fe &a) {
    if (fe_is_zero(a)) return a;
    fe r;
    u32 m[8];
    FqMod::limbs(m);
    sub8(r.l, m, a.l);
    return r;
}

// ---- redundant-range forms for the accumulate loop: values in [0, 2p), no reduction after a product
PK_HD void fq_two_p(u32 *t) {  // 2p
    t[0] = 0xb0f9fa8eu; t[1] = 0x7841182du; t[2] = 0xd0e3951au; t[3] = 0x2f02d522u;
    t[4] = 0x0302b0bbu; t[5] = 0x70a08b6du; t[6] = 0xc2634053u; t[7] = 0x60c89ce5u;
}
PK_HD fe fq_mul_lz(const fe &a, const fe &b) { return mont_mul<FqMod, true>(a, b); }
PK_HD fe fq_sqr_lz(const fe &a) { return mont_sqr<FqMod, true>(a); }
PK_HD fe fq_mul_sum_lz(const fe &a, const fe &b, const fe &c, const fe &d) { return mont_mul_sum<FqMod, true>(a, b, c, d); }
PK_HD fe fq_add_lz(const fe &a, const fe &b) {  // a + b < 4p < 2^256, then minus 2p if it reaches it
    fe r;
    u32 t[8], d[8];
    fq_two_p(t);
    add8(r.l, a.l, b.l);
    if (sub8(d, r.l, t) == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l[i] = d[i];
    }
    return r;
}
PK_HD fe fq_sub_lz(const fe &a, const fe &b) {
    fe r;
    if (sub8(r.l, a.l, b.l)) {
        u32 t[8];
        fq_two_p(t);
        add8(r.l, r.l, t);
    }
    return r;
}
PK_HD fe fq_neg_lz(const fe &a) {  // 2p - a, 0 -> 0
    if (fe_is_zero(a)) return a;
    fe r;
    u32 t[8];
    fq_two_p(t);
    sub8(r.l, t, a.l);
    return r;
}
PK_HD bool fq_is_zero_lz(const fe &a) {  // a in [0, 2p): congruent to zero iff 0 or p
    u32 m[8];
    FqMod::limbs(m);
    u32 d = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) d |= a.l[i] ^ m[i];
    return fe_is_zero(a) || d == 0;
}
PK_HD fe fq_canonical(const fe &a) {  // [0, 2p) -> [0, p)
    fe r = a;
    final_sub<FqMod>(r.l);
    return r;
}

// a^(p-2): 0 -> 0.  Only used once per MSM (projective -> affine).
PK_HD fe fq_inv(const fe &a) {
    u32 e[8];
    FqMod::limbs(e);
    e[0] -= 2;  // p ends in ...47, no borrow
    fe acc = fq_one(), base = a;
    for (int i = 0; i < 254; ++i) {
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fq_mul(acc, base);
        base = fq_sqr(base);
    }
    return acc;
}

// ------------------------------------------------------------ fast inversion
// Modular inverse by the Bernstein-Yang "safegcd" division steps, 30 steps per batch on
// signed 30-bit limbs (the formulation popularised by libsecp256k1's modinv32): 20 batches
// = 600 steps >= the 590 needed for a 256-bit modulus, constant time and branch-free, so
// every lane of a warp inverts its own value in lock step.  About 15 K instructions against
// 87 K for the Fermat ladder (380 products); used for the batched affine additions and the
// final XYZZ -> affine.  Plain C integer arithmetic: the same code runs in the CPU tests.
struct s30 {
    int32_t v[9];  // value = sum v[i] * 2^(30 i)
};
struct trans30 {
    int32_t u, v, q, r;
};

PK_HD s30 fq_modulus_s30() {
    s30 m = {{0x187cfd47, 0x3082305b, 0x071ca8d3, 0x205aa45a, 0x01585d97, 0x0116da06, 0x1a029b85, 0x139cb84c, 0x3064}};
    return m;
}
#define PK_P_INV30 0x1b799c77u  // p^-1 mod 2^30

PK_HD s30 fe_to_s30(const fe &a) {
    s30 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int bit = 30 * i, word = bit >> 5, sh = bit & 31;
        u32 val = a.l[word] >> sh;
        if (sh > 2 && word + 1 < 8) val |= a.l[word + 1] << (32 - sh);
        r.v[i] = (int32_t)(val & 0x3fffffffu);
    }
    return r;
}
// r normalised: limbs in [0, 2^30), value < 2^256.
PK_HD fe s30_to_fe(const s30 &r) {
    fe o;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const int bit = 32 * w, limb = bit / 30, sh = bit % 30;
        u32 val = (u32)r.v[limb] >> sh;
        if (limb + 1 < 9) val |= (u32)r.v[limb + 1] << (30 - sh);
        if (sh > 28 && limb + 2 < 9) val |= (u32)r.v[limb + 2] << (60 - sh);
        o.l[w] = val;
    }
    return o;
}

PK_HD int32_t divsteps_30(int32_t zeta, u32 f0, u32 g0, trans30 &t) {
    u32 u = 1, v = 0, q = 0, r = 1;
    u32 f = f0, g = g0;
#pragma unroll 1
    for (int i = 0; i < 30; ++i) {
        u32 c1 = (u32)(zeta >> 31);
        const u32 c2 = 0u - (g & 1u);
        const u32 x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
        g += x & c2; q += y & c2; r += z & c2;
        c1 &= c2;
        zeta = (int32_t)(((u32)zeta ^ c1) - 1u);
        f += g & c1; u += q & c1; v += r & c1;
        g >>= 1; u <<= 1; v <<= 1;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return zeta;
}

// [d, e] <- t / 2^30 * [d, e] mod p
PK_HD void update_de_30(s30 &d, s30 &e, const trans30 &t) {
    const int32_t M30 = 0x3fffffff;
    const s30 m = fq_modulus_s30();
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
    int32_t md = (t.u & sd) + (t.v & se);
    int32_t me = (t.q & sd) + (t.r & se);
    int32_t di = d.v[0], ei = e.v[0];
    int64_t cd = u * di + v * ei;
    int64_t ce = q * di + r * ei;
    md -= (int32_t)((PK_P_INV30 * (u32)cd + (u32)md) & (u32)M30);
    me -= (int32_t)((PK_P_INV30 * (u32)ce + (u32)me) & (u32)M30);
    cd += (int64_t)m.v[0] * md;
    ce += (int64_t)m.v[0] * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; ++i) {
        di = d.v[i]; ei = e.v[i];
        cd += u * di + v * ei;
        ce += q * di + r * ei;
        cd += (int64_t)m.v[i] * md;
        ce += (int64_t)m.v[i] * me;
        d.v[i - 1] = (int32_t)cd & M30; cd >>= 30;
        e.v[i - 1] = (int32_t)ce & M30; ce >>= 30;
    }
    d.v[8] = (int32_t)cd;
    e.v[8] = (int32_t)ce;
}

// [f, g] <- t / 2^30 * [f, g]
PK_HD void update_fg_30(s30 &f, s30 &g, const trans30 &t) {
    const int32_t M30 = 0x3fffffff;
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    int32_t fi = f.v[0], gi = g.v[0];
    int64_t cf = u * fi + v * gi;
    int64_t cg = q * fi + r * gi;
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; ++i) {
        fi = f.v[i]; gi = g.v[i];
        cf += u * fi + v * gi;
        cg += q * fi + r * gi;
        f.v[i - 1] = (int32_t)cf & M30; cf >>= 30;
        g.v[i - 1] = (int32_t)cg & M30; cg >>= 30;
    }
    f.v[8] = (int32_t)cf;
    g.v[8] = (int32_t)cg;
}

// r in (-2p, p) with limbs in (-2^30, 2^30) -> [0, p), negated first when sign < 0.
PK_HD void normalize_30(s30 &r, int32_t sign) {
    const int32_t M30 = 0x3fffffff;
    const s30 m = fq_modulus_s30();
    int32_t cond_add = r.v[8] >> 31;
    const int32_t cond_negate = sign >> 31;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        r.v[i] += m.v[i] & cond_add;
        r.v[i] = (r.v[i] ^ cond_negate) - cond_negate;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.v[i + 1] += r.v[i] >> 30;
        r.v[i] &= M30;
    }
    cond_add = r.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.v[i] += m.v[i] & cond_add;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.v[i + 1] += r.v[i] >> 30;
        r.v[i] &= M30;
    }
}

// x^-1 mod p for a plain integer 0 <= x < p (0 -> 0).
PK_HD fe fq_inv_plain(const fe &x) {
    s30 d = {{0, 0, 0, 0, 0, 0, 0, 0, 0}}, e = {{1, 0, 0, 0, 0, 0, 0, 0, 0}};
    s30 f = fq_modulus_s30(), g = fe_to_s30(x);
    int32_t zeta = -1;
#pragma unroll 1
    for (int i = 0; i < 20; ++i) {
        trans30 t;
        zeta = divsteps_30(zeta, (u32)f.v[0], (u32)g.v[0], t);
        update_de_30(d, e, t);
        update_fg_30(f, g, t);
    }
    normalize_30(d, f.v[8]);
    return s30_to_fe(d);
}

// Montgomery inverse: (aR)^-1 = a^-1 R^-1, times R^2 (one product with R^3) gives a^-1 R.
PK_HD fe fq_inv_fast(const fe &a) {
    const fe r3 = {{0xda1530dfu, 0xb1cd6dafu, 0xa7283db6u, 0x62f210e6u, 0x0ada0afbu, 0xef7f0b0cu, 0x2d592544u, 0x20fd6e90u}};
    return fq_mul(fq_inv_plain(a), r3);
}

// Fr Montgomery form -> canonical integer (halo2_curves `to_repr`, called at
// msm.rs:153): one Montgomery reduction, i.e. a product with the plain integer 1.
// a / 2^256 mod m for a < m: the reduction half of a Montgomery product (8 rounds of T = (T + q*m) >> 32 with
// q = T[0] * (-m^-1)), 64 + 8 multiply instructions instead of the 136 of a product with one.  T stays below 2m.
template <class MOD>
PK_HD fe mont_reduce(const fe &a) {
    u32 m[8], t[8];
    MOD::limbs(m);
#pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = a.l[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const u32 q = t[0] * MOD::inv();
        u64 c = ((u64)q * m[0] + t[0]) >> 32;  // low word is zero by construction
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            const u64 p = (u64)q * m[j] + t[j] + c;
            t[j - 1] = (u32)p;
            c = p >> 32;
        }
        t[7] = (u32)c;  // T < 2m < 2^255: no ninth word
    }
    fe r;
#pragma unroll
    for (int i = 0; i < 8; ++i) r.l[i] = t[i];
    final_sub<MOD>(r.l);
    return r;
}

PK_HD fe fr_to_canonical(const fe &a) { return mont_reduce<FrMod>(a); }

}  // namespace pk
