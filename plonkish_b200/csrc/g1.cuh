// BN254 G1 point arithmetic on the device: extended Jacobian ("XYZZ")
// accumulators (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), affine inputs.
//
// Replaces the halo2_curves 0.3.3 [ext] group law that the reference reaches at
// /root/reference/plonkish_backend/src/util/arithmetic/msm.rs:133 (affine +
// affine), :135 (projective += affine), :145-148, :163 (double), :178.  The
// projective representation differs from the reference's Jacobian (x, y, z); the
// affine value every caller takes (`.into()` / `.to_affine()`, e.g.
// pcs/multilinear/kzg.rs:255) is the same field elements bit for bit.
//
// All exceptional cases are handled (identity operands, P + P, P + (-P)): MSM
// inputs from the reference include identity bases and repeated bases
// (accumulation/protostar.rs:207-212).
#pragma once
#include "fq.cuh"

namespace pk {

// The point formulas are templated on how a field product is issued.  MulInline expands
// the 270-instruction product in place: right for the accumulate loop, which is bound by
// the integer pipe (measured: 37.6 ms inlined vs 39.8 ms with calls at 2^24).  MulCall
// routes through one out-of-line copy: right for the latency-bound tail kernels (bucket
// reduce 1.45 -> 1.04 ms, item levels 0.43 -> 0.35 ms) whose inlined bodies (up to 35 K
// instructions) thrash the instruction cache.
struct MulInline {
    static PK_HD fe mul(const fe &a, const fe &b) { return fq_mul(a, b); }
    static PK_HD fe mul_sum(const fe &a, const fe &b, const fe &c, const fe &d) { return fq_mul_sum(a, b, c, d); }
    static PK_HD fe sqr(const fe &a) { return fq_sqr(a); }
};
#if !defined(PLONKISH_EMUL)
__device__ __noinline__ fe fq_mul_call(fe a, fe b) { return mont_mul<FqMod>(a, b); }
__device__ __noinline__ fe fq_mul_sum_call(fe a, fe b, fe c, fe d) { return mont_mul_sum<FqMod>(a, b, c, d); }
__device__ __noinline__ fe fq_sqr_call(fe a) { return mont_sqr<FqMod>(a); }
struct MulCall {
    static PK_HD fe mul(const fe &a, const fe &b) { return fq_mul_call(a, b); }
    static PK_HD fe mul_sum(const fe &a, const fe &b, const fe &c, const fe &d) { return fq_mul_sum_call(a, b, c, d); }
    static PK_HD fe sqr(const fe &a) { return fq_sqr_call(a); }
};
#else
typedef MulInline MulCall;
#endif
#define PK_MUL(a, b) M::mul(a, b)
#define PK_SQR(a) M::sqr(a)
// a*b - c*d with one Montgomery reduction (the y-coordinate of every formula below)
#define PK_MUL_DIFF(a, b, c, d) M::mul_sum(a, b, fq_neg(c), d)

struct affine {  // 64 B, (0,0) = identity — the layout of halo2_curves G1Affine [ext]
    fe x, y;
};
struct xyzz {  // 128 B; identity <=> zz == 0
    fe x, y, zz, zzz;
};

PK_HD bool affine_is_identity(const affine &p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }
PK_HD bool xyzz_is_identity(const xyzz &p) { return fe_is_zero(p.zz); }
PK_HD xyzz xyzz_identity() {
    xyzz r;
    r.x = fe_zero(); r.y = fe_zero(); r.zz = fe_zero(); r.zzz = fe_zero();
    return r;
}
PK_HD xyzz xyzz_from_affine(const affine &p) {
    xyzz r;
    if (affine_is_identity(p)) return xyzz_identity();
    r.x = p.x; r.y = p.y; r.zz = fq_one(); r.zzz = fq_one();
    return r;
}

// 2 * (affine P), P != identity  (mdbl-2008-s-1 with a = 0): 3M + 2S... spelled out:
// U = 2y, V = U^2, W = U*V, S = x*V, M = 3x^2, X3 = M^2 - 2S, Y3 = M(S - X3) - W*y.
template <class M = MulCall>
PK_HD xyzz xyzz_double_affine(const affine &p) {
    xyzz r;
    fe u = fq_dbl(p.y);
    fe v = PK_SQR(u);
    fe w = PK_MUL(u, v);
    fe s = PK_MUL(p.x, v);
    fe xx = PK_SQR(p.x);
    fe m = fq_add(fq_dbl(xx), xx);
    r.x = fq_sub(PK_SQR(m), fq_dbl(s));
    r.y = PK_MUL_DIFF(m, fq_sub(s, r.x), w, p.y);
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * P for an XYZZ point (dbl-2008-s-1, a = 0).
template <class M = MulCall>
PK_HD xyzz xyzz_double(const xyzz &p) {
    if (xyzz_is_identity(p)) return p;
    xyzz r;
    fe u = fq_dbl(p.y);
    fe v = PK_SQR(u);
    fe w = PK_MUL(u, v);
    fe s = PK_MUL(p.x, v);
    fe xx = PK_SQR(p.x);
    fe m = fq_add(fq_dbl(xx), xx);
    r.x = fq_sub(PK_SQR(m), fq_dbl(s));
    r.y = PK_MUL_DIFF(m, fq_sub(s, r.x), w, p.y);
    r.zz = PK_MUL(v, p.zz);
    r.zzz = PK_MUL(w, p.zzz);
    return r;
}

// acc += (x2, y2) with the affine point given as two field elements (so a caller
// can pass a negated y without building a struct).  madd-2008-s: 8M + 2S.
template <class M = MulCall>
PK_HD void xyzz_madd(xyzz &acc, const fe &x2, const fe &y2) {
    if (fe_is_zero(x2) && fe_is_zero(y2)) return;  // identity base
    if (xyzz_is_identity(acc)) {
        acc.x = x2; acc.y = y2; acc.zz = fq_one(); acc.zzz = fq_one();
        return;
    }
    fe u2 = PK_MUL(x2, acc.zz);
    fe s2 = PK_MUL(y2, acc.zzz);
    fe p = fq_sub(u2, acc.x);
    fe r = fq_sub(s2, acc.y);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) {
            affine q; q.x = x2; q.y = y2;
            acc = xyzz_double_affine<M>(q);
        } else {
            acc = xyzz_identity();
        }
        return;
    }
    fe pp = PK_SQR(p);
    fe ppp = PK_MUL(p, pp);
    fe q = PK_MUL(acc.x, pp);
    fe x3 = fq_sub(fq_sub(PK_SQR(r), ppp), fq_dbl(q));
    fe y3 = PK_MUL_DIFF(r, fq_sub(q, x3), acc.y, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = PK_MUL(acc.zz, pp);
    acc.zzz = PK_MUL(acc.zzz, ppp);
}

// The accumulate loop's form of the mixed addition: the accumulator's coordinates live in the redundant range
// [0, 2p) between additions (fq_*_lz: no conditional subtraction after a product, the y-coordinate's fused pair keeps
// one), which takes ~150 ALU instructions out of every addition; xyzz_canonical brings a sum back to [0, p) before it
// is stored.  The affine input (x2, y2) is canonical, as it comes from memory.
PK_HD xyzz xyzz_canonical(const xyzz &p) {
    xyzz r;
    r.x = fq_canonical(p.x); r.y = fq_canonical(p.y); r.zz = fq_canonical(p.zz); r.zzz = fq_canonical(p.zzz);
    return r;
}
PK_HD void xyzz_madd_lazy(xyzz &acc, const fe &x2, const fe &y2) {
    if (fe_is_zero(x2) && fe_is_zero(y2)) return;  // identity base
    if (xyzz_is_identity(acc)) {
        acc.x = x2; acc.y = y2; acc.zz = fq_one(); acc.zzz = fq_one();
        return;
    }
    const fe u2 = fq_mul_lz(x2, acc.zz);
    const fe s2 = fq_mul_lz(y2, acc.zzz);
    const fe p = fq_sub_lz(u2, acc.x);
    const fe r = fq_sub_lz(s2, acc.y);
    if (fq_is_zero_lz(p)) {
        if (fq_is_zero_lz(r)) {
            affine q; q.x = x2; q.y = y2;
            acc = xyzz_double_affine<MulInline>(q);
        } else {
            acc = xyzz_identity();
        }
        return;
    }
    const fe pp = fq_sqr_lz(p);
    const fe ppp = fq_mul_lz(p, pp);
    const fe q = fq_mul_lz(acc.x, pp);
    const fe x3 = fq_sub_lz(fq_sub_lz(fq_sqr_lz(r), ppp), fq_add_lz(q, q));
    const fe y3 = fq_mul_sum_lz(r, fq_sub_lz(q, x3), fq_neg_lz(acc.y), ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = fq_mul_lz(acc.zz, pp);
    acc.zzz = fq_mul_lz(acc.zzz, ppp);
}

// a + b, both XYZZ (add-2008-s): 12M + 2S.
template <class M = MulCall>
PK_HD xyzz xyzz_add(const xyzz &a, const xyzz &b) {
    if (xyzz_is_identity(a)) return b;
    if (xyzz_is_identity(b)) return a;
    fe u1 = PK_MUL(a.x, b.zz);
    fe u2 = PK_MUL(b.x, a.zz);
    fe s1 = PK_MUL(a.y, b.zzz);
    fe s2 = PK_MUL(b.y, a.zzz);
    fe p = fq_sub(u2, u1);
    fe r = fq_sub(s2, s1);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) return xyzz_double<M>(a);
        return xyzz_identity();
    }
    fe pp = PK_SQR(p);
    fe ppp = PK_MUL(p, pp);
    fe q = PK_MUL(u1, pp);
    xyzz o;
    o.x = fq_sub(fq_sub(PK_SQR(r), ppp), fq_dbl(q));
    o.y = PK_MUL_DIFF(r, fq_sub(q, o.x), s1, ppp);
    o.zz = PK_MUL(PK_MUL(a.zz, b.zz), pp);
    o.zzz = PK_MUL(PK_MUL(a.zzz, b.zzz), ppp);
    return o;
}

// XYZZ -> affine with one inversion: I = (ZZ*ZZZ)^-1, 1/ZZ = I*ZZZ, 1/ZZZ = I*ZZ.
template <class M = MulCall>
PK_HD affine xyzz_to_affine(const xyzz &p) {
    affine r;
    if (xyzz_is_identity(p)) {
        r.x = fe_zero(); r.y = fe_zero();
        return r;
    }
    fe i = fq_inv_fast(PK_MUL(p.zz, p.zzz));
    r.x = PK_MUL(p.x, PK_MUL(i, p.zzz));
    r.y = PK_MUL(p.y, PK_MUL(i, p.zz));
    return r;
}

#undef PK_MUL
#undef PK_SQR
#undef PK_MUL_DIFF

}  // namespace pk
