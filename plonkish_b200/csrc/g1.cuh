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
PK_HD xyzz xyzz_double_affine(const affine &p) {
    xyzz r;
    fe u = fq_dbl(p.y);
    fe v = fq_sqr(u);
    fe w = fq_mul(u, v);
    fe s = fq_mul(p.x, v);
    fe xx = fq_sqr(p.x);
    fe m = fq_add(fq_dbl(xx), xx);
    r.x = fq_sub(fq_sqr(m), fq_dbl(s));
    r.y = fq_sub(fq_mul(m, fq_sub(s, r.x)), fq_mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * P for an XYZZ point (dbl-2008-s-1, a = 0).
PK_HD xyzz xyzz_double(const xyzz &p) {
    if (xyzz_is_identity(p)) return p;
    xyzz r;
    fe u = fq_dbl(p.y);
    fe v = fq_sqr(u);
    fe w = fq_mul(u, v);
    fe s = fq_mul(p.x, v);
    fe xx = fq_sqr(p.x);
    fe m = fq_add(fq_dbl(xx), xx);
    r.x = fq_sub(fq_sqr(m), fq_dbl(s));
    r.y = fq_sub(fq_mul(m, fq_sub(s, r.x)), fq_mul(w, p.y));
    r.zz = fq_mul(v, p.zz);
    r.zzz = fq_mul(w, p.zzz);
    return r;
}

// acc += (x2, y2) with the affine point given as two field elements (so a caller
// can pass a negated y without building a struct).  madd-2008-s: 8M + 2S.
PK_HD void xyzz_madd(xyzz &acc, const fe &x2, const fe &y2) {
    if (fe_is_zero(x2) && fe_is_zero(y2)) return;  // identity base
    if (xyzz_is_identity(acc)) {
        acc.x = x2; acc.y = y2; acc.zz = fq_one(); acc.zzz = fq_one();
        return;
    }
    fe u2 = fq_mul(x2, acc.zz);
    fe s2 = fq_mul(y2, acc.zzz);
    fe p = fq_sub(u2, acc.x);
    fe r = fq_sub(s2, acc.y);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) {
            affine q; q.x = x2; q.y = y2;
            acc = xyzz_double_affine(q);
        } else {
            acc = xyzz_identity();
        }
        return;
    }
    fe pp = fq_sqr(p);
    fe ppp = fq_mul(p, pp);
    fe q = fq_mul(acc.x, pp);
    fe x3 = fq_sub(fq_sub(fq_sqr(r), ppp), fq_dbl(q));
    fe y3 = fq_sub(fq_mul(r, fq_sub(q, x3)), fq_mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = fq_mul(acc.zz, pp);
    acc.zzz = fq_mul(acc.zzz, ppp);
}

// a + b, both XYZZ (add-2008-s): 12M + 2S.
PK_HD xyzz xyzz_add(const xyzz &a, const xyzz &b) {
    if (xyzz_is_identity(a)) return b;
    if (xyzz_is_identity(b)) return a;
    fe u1 = fq_mul(a.x, b.zz);
    fe u2 = fq_mul(b.x, a.zz);
    fe s1 = fq_mul(a.y, b.zzz);
    fe s2 = fq_mul(b.y, a.zzz);
    fe p = fq_sub(u2, u1);
    fe r = fq_sub(s2, s1);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r)) return xyzz_double(a);
        return xyzz_identity();
    }
    fe pp = fq_sqr(p);
    fe ppp = fq_mul(p, pp);
    fe q = fq_mul(u1, pp);
    xyzz o;
    o.x = fq_sub(fq_sub(fq_sqr(r), ppp), fq_dbl(q));
    o.y = fq_sub(fq_mul(r, fq_sub(q, o.x)), fq_mul(s1, ppp));
    o.zz = fq_mul(fq_mul(a.zz, b.zz), pp);
    o.zzz = fq_mul(fq_mul(a.zzz, b.zzz), ppp);
    return o;
}

// XYZZ -> affine with one inversion: I = (ZZ*ZZZ)^-1, 1/ZZ = I*ZZZ, 1/ZZZ = I*ZZ.
PK_HD affine xyzz_to_affine(const xyzz &p) {
    affine r;
    if (xyzz_is_identity(p)) {
        r.x = fe_zero(); r.y = fe_zero();
        return r;
    }
    fe i = fq_inv(fq_mul(p.zz, p.zzz));
    r.x = fq_mul(p.x, fq_mul(i, p.zzz));
    r.y = fq_mul(p.y, fq_mul(i, p.zz));
    return r;
}

}  // namespace pk
