// BN254 Fq arithmetic on the FP64 pipe of the B200: 254-bit values as six 48-bit limbs held in doubles,
// Montgomery products built from round-toward-zero DFMAs.
//
// Why: the integer multiplier (IMAD.WIDE, fq.cuh) is the roof of the bucket accumulation — and it is one pipe.
// The B200 keeps a full-rate FP64 pipe next to it (measured: 16.8 T DFMA/s alone; interleaved 1:1 with
// mad.wide.u32 both streams run at 9.6 T/s, i.e. the integer pipe keeps its full rate while the DFMAs ride along —
// plonkish_cuda_bench_fp64_pipe).  A second accumulate kernel written entirely in this arithmetic runs on the same
// SMs at the same time and takes its share of the sorted entries (msm_kernels.cuh, k_accumulate_dp).
//
// How a product is formed: for integers 0 <= a, b < 2^48 held exactly in doubles,
//     A' = fma_rz(a, b, A)          with A a multiple of 2^48 in [2^100, 2^101): the ulp there is 2^48, so the
//                                   truncated sum is A + floor(a*b / 2^48) * 2^48          (exact high part)
//     L  = fma_rz(a, b, -(A' - A))  = a*b mod 2^48                                        (exact low part)
// i.e. the high parts of a whole column of partial products accumulate inside the FMA chain itself and the low parts
// in one more DADD each: four FP64 instructions per 48x48-bit partial product, no integer instruction at all.
// A Montgomery product (product scanning, R = 2^288) is 72 such partial products plus 73 bookkeeping operations.
//
// Values: dfe = 6 doubles, each an integer in [0, 2^48) ("normalised"), value = sum l[i] * 2^(48 i).  The Montgomery
// radix here is 2^288, so x is held as x * 2^288 mod p; memory keeps the library's usual x * 2^256 mod p
// (dp_from_mont256 / dp_to_mont256 convert with one product each).  Products return values in [0, p + 2^240);
// additions and subtractions return normalised limbs and values below a few dozen p; dp_canonical brings a value
// into [0, p).
//
// The same source compiles under g++ for the CPU test suite (tests/emul): fma() under FE_TOWARDZERO is __fma_rz.
// Reference: replaces, for the share of the work it takes, the same halo2_curves 0.3.3 [ext] Fq arithmetic as
// fq.cuh (/root/reference/plonkish_backend/src/util/arithmetic/msm.rs:133-148).
#pragma once
#include "fq.cuh"

#ifdef PLONKISH_EMUL
#include <cfenv>
#include <cmath>
#endif

namespace pk {

struct dfe {
    double l[6];
};

#define DP_C100 1267650600228229401496703205376.0 /* 2^100 */
#define DP_2P48 281474976710656.0
#define DP_2M48 (1.0 / 281474976710656.0)
#define DP_2P52 4503599627370496.0

#ifndef PLONKISH_EMUL
PK_HD double dp_fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
PK_HD double dp_add_rz(double a, double b) { return __dadd_rz(a, b); }
PK_HD double dp_add(double a, double b) { return __dadd_rn(a, b); }   // exact in every use below: the intrinsic only stops contraction
PK_HD double dp_sub(double a, double b) { return __dsub_rn(a, b); }
PK_HD double dp_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
PK_HD double dp_from_bits(u32 hi, u32 lo) { return __hiloint2double((int)hi, (int)lo); }
PK_HD u32 dp_lo_bits(double v) { return (u32)__double2loint(v); }
PK_HD u32 dp_hi_bits(double v) { return (u32)__double2hiint(v); }
struct DpRoundGuard {};
#else
// The emulation build evaluates these under FE_TOWARDZERO (DpRoundGuard in every entry function); volatile keeps the
// compiler from folding or contracting across the rounding-mode change.
inline double dp_fma_rz(double a, double b, double c) { volatile double r = fma(a, b, c); return r; }
inline double dp_add_rz(double a, double b) { volatile double r = a + b; return r; }
inline double dp_add(double a, double b) { volatile double r = a + b; return r; }
inline double dp_sub(double a, double b) { volatile double r = a - b; return r; }
inline double dp_fma(double a, double b, double c) { volatile double r = fma(a, b, c); return r; }
inline double dp_from_bits(u32 hi, u32 lo) { u64 b = ((u64)hi << 32) | lo; double d; memcpy(&d, &b, 8); return d; }
inline u32 dp_lo_bits(double v) { u64 b; memcpy(&b, &v, 8); return (u32)b; }
inline u32 dp_hi_bits(double v) { u64 b; memcpy(&b, &v, 8); return (u32)(b >> 32); }
struct DpRoundGuard {
    int saved;
    DpRoundGuard() : saved(fegetround()) { fesetround(FE_TOWARDZERO); }
    ~DpRoundGuard() { fesetround(saved); }
};
#endif

// p in 48-bit limbs, -p^-1 mod 2^48.
#define DP_P0 154029749239111.0 /* 0x8c16d87cfd47 */
#define DP_P1 114837938846752.0 /* 0x6871ca8d3c20 */
#define DP_P2 97158997043857.0  /* 0x585d97816a91 */
#define DP_P3 202654906483073.0 /* 0xb85045b68181 */
#define DP_P4 86255311364137.0  /* 0x4e72e131a029 */
#define DP_P5 12388.0           /* 0x3064 */
#define DP_PINV 8258761155465.0 /* 0x782e4866389 */

PK_HD double dp_p(int i) {
    switch (i) {
        case 0: return DP_P0;
        case 1: return DP_P1;
        case 2: return DP_P2;
        case 3: return DP_P3;
        case 4: return DP_P4;
        default: return DP_P5;
    }
}

// One 48x48-bit partial product into a column: A (multiple of 2^48 in [2^100, 2^101)) takes the high part, B the low.
PK_HD void dp_mac(double a, double b, double &A, double &B) {
    const double An = dp_fma_rz(a, b, A);
    const double d = dp_sub(An, A);
    B = dp_add(B, dp_fma_rz(a, b, -d));
    A = An;
}
// v an integer in [0, 2^53): lo = v mod 2^48, hi = v - lo.
PK_HD void dp_split48(double v, double &lo, double &hi) {
    hi = dp_sub(dp_add_rz(v, DP_C100), DP_C100);
    lo = dp_sub(v, hi);
}

// a * b / 2^288 mod p for normalised limbs (any values below 2^270); result normalised, in [0, p + 2^240).
// Product scanning: column k collects a_i b_(k-i) and q_i p_(k-i); q_k clears the column's low 48 bits.
// Bounds: a column takes at most 13 partial products, so the high accumulator stays below 2^100 + 13 * 2^96 and the
// low one (12-13 low parts + the carry of the previous column, itself < 13 * 2^48) below 26 * 2^48 < 2^53.
PK_HD dfe dp_mul(const dfe &a, const dfe &b) {
    double q[6];
    dfe r;
    double carry = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double A = DP_C100, B = carry;
#pragma unroll
        for (int i = 0; i <= k; ++i) dp_mac(a.l[i], b.l[k - i], A, B);
#pragma unroll
        for (int i = 0; i < k; ++i) dp_mac(q[i], dp_p(k - i), A, B);
        double lo, hi;
        dp_split48(B, lo, hi);
        const double Aq = dp_fma_rz(lo, DP_PINV, DP_C100);
        q[k] = dp_fma_rz(lo, DP_PINV, -dp_sub(Aq, DP_C100));  // lo * (-p^-1) mod 2^48
        dp_mac(q[k], DP_P0, A, B);                             // B is a multiple of 2^48 now
        carry = dp_fma(dp_add(A, B), DP_2M48, -DP_2P52);       // (A - 2^100 + B) / 2^48, exact
    }
#pragma unroll
    for (int k = 6; k < 11; ++k) {
        double A = DP_C100, B = carry;
#pragma unroll
        for (int i = k - 5; i <= 5; ++i) {
            dp_mac(a.l[i], b.l[k - i], A, B);
            dp_mac(q[i], dp_p(k - i), A, B);
        }
        double lo, hi;
        dp_split48(B, lo, hi);
        r.l[k - 6] = lo;
        carry = dp_fma(dp_add(A, hi), DP_2M48, -DP_2P52);
    }
    r.l[5] = carry;
    return r;
}

// a * a / 2^288 mod p: the symmetric partial products are taken once against the doubled limb (2 a_j < 2^49, so a
// column's high parts stay below 2^100 + 13 * 2^97 < 2^101 and its low parts below 2^53 as before).
PK_HD dfe dp_sqr(const dfe &a) {
    double q[6], a2[6];
    dfe r;
    double carry = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) a2[i] = dp_add(a.l[i], a.l[i]);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double A = DP_C100, B = carry;
#pragma unroll
        for (int i = 0; 2 * i < k; ++i) dp_mac(a.l[i], a2[k - i], A, B);
        if ((k & 1) == 0) dp_mac(a.l[k / 2], a.l[k / 2], A, B);
#pragma unroll
        for (int i = 0; i < k; ++i) dp_mac(q[i], dp_p(k - i), A, B);
        double lo, hi;
        dp_split48(B, lo, hi);
        const double Aq = dp_fma_rz(lo, DP_PINV, DP_C100);
        q[k] = dp_fma_rz(lo, DP_PINV, -dp_sub(Aq, DP_C100));
        dp_mac(q[k], DP_P0, A, B);
        carry = dp_fma(dp_add(A, B), DP_2M48, -DP_2P52);
    }
#pragma unroll
    for (int k = 6; k < 11; ++k) {
        double A = DP_C100, B = carry;
#pragma unroll
        for (int i = k - 5; 2 * i < k; ++i) dp_mac(a.l[i], a2[k - i], A, B);
        if ((k & 1) == 0) dp_mac(a.l[k / 2], a.l[k / 2], A, B);
#pragma unroll
        for (int i = k - 5; i <= 5; ++i) dp_mac(q[i], dp_p(k - i), A, B);
        double lo, hi;
        dp_split48(B, lo, hi);
        r.l[k - 6] = lo;
        carry = dp_fma(dp_add(A, hi), DP_2M48, -DP_2P52);
    }
    r.l[5] = carry;
    return r;
}

// Carry propagation: limbs in [0, 2^52) on entry (integers), normalised on exit; the top limb keeps what is left.
PK_HD dfe dp_normalize(const dfe &a) {
    dfe r;
    double carry = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double lo, hi;
        dp_split48(dp_add(a.l[i], carry), lo, hi);
        r.l[i] = lo;
        carry = hi * DP_2M48;
    }
    r.l[5] = dp_add(a.l[5], carry);
    return r;
}
PK_HD dfe dp_add_fe(const dfe &a, const dfe &b) {
    dfe t;
#pragma unroll
    for (int i = 0; i < 6; ++i) t.l[i] = dp_add(a.l[i], b.l[i]);
    return dp_normalize(t);
}
// a - b + K p for K = 2, 4, 8 and b < (K - 1) p + 2^241: K p is spelled with every limb below the top one raised by
// 2^48 (borrowed from the limb above), so no limb difference goes negative; the result is normalised and
// below a + K p.
template <int K>
PK_HD dfe dp_sub_fe(const dfe &a, const dfe &b) {
    const double m2[6] = {308059498478222.0 /* 0x1182db0f9fa8e */, 511150854404160.0 /* 0x1d0e3951a7840 */, 475792970798369.0 /* 0x1b0bb2f02d521 */,
                          405309812966145.0 /* 0x170a08b6d0301 */, 453985599438930.0 /* 0x19ce5c2634052 */, 24775.0 /* 0x60c7 */};
    const double m4[6] = {334644020245788.0 /* 0x1305b61f3f51c */, 459351755387009.0 /* 0x1a1c72a34f081 */, 388635988175428.0 /* 0x161765e05aa44 */,
                          529144649221636.0 /* 0x1e14116da0604 */, 345021245456549.0 /* 0x139cb84c680a5 */, 49552.0 /* 0xc190 */};
    const double m8[6] = {387813063780920.0 /* 0x160b6c3e7ea38 */, 355753557352707.0 /* 0x1438e5469e103 */, 495796999640202.0 /* 0x1c2ecbc0b548a */,
                          495339345021961.0 /* 0x1c2822db40c09 */, 408567514202444.0 /* 0x17397098d014c */, 99105.0 /* 0x18321 */};
    dfe t;
#pragma unroll
    for (int i = 0; i < 6; ++i) t.l[i] = dp_add(dp_sub(a.l[i], b.l[i]), K == 2 ? m2[i] : (K == 4 ? m4[i] : m8[i]));
    return dp_normalize(t);
}
PK_HD dfe dp_zero() {
    dfe r = {{0, 0, 0, 0, 0, 0}};
    return r;
}
PK_HD bool dp_is_zero_limbs(const dfe &a) { return (a.l[0] == 0.0) & (a.l[1] == 0.0) & (a.l[2] == 0.0) & (a.l[3] == 0.0) & (a.l[4] == 0.0) & (a.l[5] == 0.0); }
// 2^288 mod p: the Montgomery one of this radix.
PK_HD dfe dp_one() {
    dfe r = {{236557864795510.0 /* 0xd725eb7ffd76 */, 57765139611009.0 /* 0x34897ea07d81 */, 220059073551101.0 /* 0xc8247ee89afd */,
              20279615087451.0 /* 0x1271b740e35b */, 3843667272302.0 /* 0x37eec6c226e */, 1724.0 /* 0x6bc */}};
    return r;
}
// Is a product (a value in [0, p + 2^240)) congruent to 0?  Then it is exactly 0 or p.
PK_HD bool dp_product_is_zero(const dfe &a) {
    const bool z = dp_is_zero_limbs(a);
    const bool isp = (a.l[0] == DP_P0) & (a.l[1] == DP_P1) & (a.l[2] == DP_P2) & (a.l[3] == DP_P3) & (a.l[4] == DP_P4) & (a.l[5] == DP_P5);
    return z | isp;
}

// ---- 8 x 32-bit words <-> 6 x 48-bit double limbs (plain integers: no Montgomery factor involved)
PK_HD double dp_limb(u32 lo32, u32 hi16) { return dp_sub(dp_from_bits(0x43300000u | hi16, lo32), DP_2P52); }
PK_HD dfe dp_from_words(const fe &w) {
    dfe r;
    r.l[0] = dp_limb(w.l[0], w.l[1] & 0xffffu);
    r.l[1] = dp_limb((w.l[1] >> 16) | (w.l[2] << 16), w.l[2] >> 16);
    r.l[2] = dp_limb(w.l[3], w.l[4] & 0xffffu);
    r.l[3] = dp_limb((w.l[4] >> 16) | (w.l[5] << 16), w.l[5] >> 16);
    r.l[4] = dp_limb(w.l[6], w.l[7] & 0xffffu);
    r.l[5] = dp_limb(w.l[7] >> 16, 0u);
    return r;
}
// normalised limbs, value < 2^256
PK_HD fe dp_to_words(const dfe &a) {
    u32 lo[6], hi[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const double t = dp_add(a.l[i], DP_2P52);
        lo[i] = dp_lo_bits(t);
        hi[i] = dp_hi_bits(t) & 0xffffu;
    }
    fe r;
    r.l[0] = lo[0];
    r.l[1] = hi[0] | (lo[1] << 16);
    r.l[2] = (lo[1] >> 16) | (hi[1] << 16);
    r.l[3] = lo[2];
    r.l[4] = hi[2] | (lo[3] << 16);
    r.l[5] = (lo[3] >> 16) | (hi[3] << 16);
    r.l[6] = lo[4];
    r.l[7] = hi[4] | (lo[5] << 16);
    return r;
}
// memory form (x * 2^256 mod p, 8 words) -> this radix (x * 2^288 mod p): one product with 2^320 mod p.
PK_HD dfe dp_from_mont256(const fe &w) {
    const dfe k = {{65144752752170.0 /* 0x3b3fb1d8c62a */, 258644892921246.0 /* 0xeb3c74f72d9e */, 15543954691923.0 /* 0x0e231be5d753 */,
                    91457802165636.0 /* 0x532e2dcf5d84 */, 2940345067620.0 /* 0x02ac9a392864 */, 8539.0 /* 0x215b */}};
    return dp_mul(dp_from_words(w), k);
}
// value in [0, 2 p) with normalised limbs -> [0, p), as words
PK_HD fe dp_canonical_words(const dfe &a) {
    fe w = dp_to_words(a);
    u32 m[8], t[8];
    FqMod::limbs(m);
    if (sub8(t, w.l, m) == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w.l[i] = t[i];
    }
    return w;
}
// this radix -> memory form, canonical: one product with 2^256 mod p, then the conditional subtraction.
PK_HD fe dp_to_mont256(const dfe &a) {
    const dfe k = {{74276183936413.0 /* 0x438dc58f0d9d */, 270235235898205.0 /* 0xf5c70b3dd35d */, 77154968202024.0 /* 0x462c0a78eb28 */,
                    112625374427257.0 /* 0x666ea36f7879 */, 131673396600623.0 /* 0x77c19a07df2f */, 3594.0 /* 0x0e0a */}};
    return dp_canonical_words(dp_mul(a, k));
}

// ---------------------------------------------------------------- points
// XYZZ accumulator in this arithmetic (the same madd-2008-s formulas as g1.cuh xyzz_madd: 8 M + 2 S).  identity <=>
// every limb of zz is zero.  Coordinate ranges between calls: x < 7 p, y < 3 p, zz and zzz products (< p + 2^240).
struct dxyzz {
    dfe x, y, zz, zzz;
};
PK_HD dxyzz dxyzz_identity() {
    dxyzz r;
    r.x = dp_zero(); r.y = dp_zero(); r.zz = dp_zero(); r.zzz = dp_zero();
    return r;
}
PK_HD bool dxyzz_is_identity(const dxyzz &p) { return dp_is_zero_limbs(p.zz); }

} // namespace pk
#include "g1.cuh"
namespace pk {

PK_HD dxyzz dxyzz_from_words(const xyzz &w) {
    dxyzz r;
    if (xyzz_is_identity(w)) return dxyzz_identity();
    r.x = dp_from_mont256(w.x); r.y = dp_from_mont256(w.y); r.zz = dp_from_mont256(w.zz); r.zzz = dp_from_mont256(w.zzz);
    return r;
}
// canonical memory form; the identity comes back as all-zero words
PK_HD xyzz dxyzz_to_words(const dxyzz &p) {
    xyzz r;
    if (dxyzz_is_identity(p)) return xyzz_identity();
    r.x = dp_to_mont256(p.x); r.y = dp_to_mont256(p.y); r.zz = dp_to_mont256(p.zz); r.zzz = dp_to_mont256(p.zzz);
    return r;
}

// acc += (x2, y2), the affine point given in memory form (x * 2^256 mod p words, (0, 0) = identity).
// P + P and P + (-P) — detected on pp = (u2 - x1)^2, a product, hence exactly 0 or p when it vanishes — and nothing
// else take the general integer-pipe routine through a conversion; random inputs never do.
PK_HD void dxyzz_madd(dxyzz &acc, const fe &x2w, const fe &y2w) {
    if (fe_is_zero(x2w) && fe_is_zero(y2w)) return;  // identity base
    const dfe x2 = dp_from_mont256(x2w), y2 = dp_from_mont256(y2w);
    if (dxyzz_is_identity(acc)) {
        acc.x = x2; acc.y = y2; acc.zz = dp_one(); acc.zzz = dp_one();
        return;
    }
    const dfe u2 = dp_mul(x2, acc.zz);
    const dfe s2 = dp_mul(y2, acc.zzz);
    const dfe p = dp_sub_fe<8>(u2, acc.x);   // acc.x < 7 p
    const dfe r = dp_sub_fe<4>(s2, acc.y);   // acc.y < 3 p
    const dfe pp = dp_sqr(p);
    if (dp_product_is_zero(pp)) {
        xyzz w = dxyzz_to_words(acc);
        xyzz_madd<MulCall>(w, x2w, y2w);
        acc = dxyzz_from_words(w);
        return;
    }
    const dfe ppp = dp_mul(p, pp);
    const dfe q = dp_mul(acc.x, pp);
    const dfe x3 = dp_sub_fe<4>(dp_sub_fe<2>(dp_sqr(r), ppp), dp_add_fe(q, q));  // < 7 p
    const dfe y3 = dp_sub_fe<2>(dp_mul(r, dp_sub_fe<8>(q, x3)), dp_mul(acc.y, ppp));  // < 3 p
    acc.x = x3;
    acc.y = y3;
    acc.zz = dp_mul(acc.zz, pp);
    acc.zzz = dp_mul(acc.zzz, ppp);
}

}  // namespace pk
