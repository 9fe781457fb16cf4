/*
 * oracle/bn254_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See bn254_oracle.h for scope, citations and the "parity unpinned" note.
 * Build: make -C oracle   (gcc -O3 -march=x86-64-v3 -shared -fPIC -pthread)
 */
#include "bn254_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

typedef struct {
    uint64_t m[4];   /* modulus                */
    uint64_t r[4];   /* R   = 2^256 mod m      */
    uint64_t r2[4];  /* R^2 mod m              */
    uint64_t inv;    /* -m^-1 mod 2^64         */
} field_params_t;

/* BN254 base field Fq and scalar field Fr (recomputed in
 * tests/test_oracle.py from p and r with Python integers). */
static const field_params_t FIELDS[2] = {
    { { 0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL },
      { 0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL },
      { 0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL },
      0x87d20782e4866389ULL },
    { { 0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL },
      { 0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL },
      { 0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL },
      0xc2e1f593efffffffULL },
};

/* ------------------------------------------------------------------ fields */

static inline int ge4(const uint64_t a[4], const uint64_t b[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return 1;
}

static inline uint64_t sub4(uint64_t out[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - b[i] - borrow;
        out[i] = (uint64_t)t;
        borrow = (uint64_t)(t >> 64) & 1;
    }
    return borrow;
}

static inline uint64_t add4(uint64_t out[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t carry = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] + b[i] + carry;
        out[i] = (uint64_t)t;
        carry = (uint64_t)(t >> 64);
    }
    return carry;
}

/* Montgomery product a*b*R^-1 mod m, coarsely-integrated operand scanning. */
static inline __attribute__((always_inline)) void mont_mul(const field_params_t *f, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
    uint64_t t[6] = { 0, 0, 0, 0, 0, 0 };
#pragma GCC unroll 4
    for (int i = 0; i < 4; ++i) {
        uint64_t carry = 0;
#pragma GCC unroll 4
        for (int j = 0; j < 4; ++j) {
            u128 cur = (u128)a[j] * b[i] + t[j] + carry;
            t[j] = (uint64_t)cur;
            carry = (uint64_t)(cur >> 64);
        }
        u128 cur = (u128)t[4] + carry;
        t[4] = (uint64_t)cur;
        t[5] = (uint64_t)(cur >> 64);

        uint64_t q = t[0] * f->inv;
        cur = (u128)q * f->m[0] + t[0];
        carry = (uint64_t)(cur >> 64);
#pragma GCC unroll 4
        for (int j = 1; j < 4; ++j) {
            cur = (u128)q * f->m[j] + t[j] + carry;
            t[j - 1] = (uint64_t)cur;
            carry = (uint64_t)(cur >> 64);
        }
        cur = (u128)t[4] + carry;
        t[3] = (uint64_t)cur;
        t[4] = t[5] + (uint64_t)(cur >> 64);
    }
    if (t[4] || ge4(t, f->m)) {
        sub4(out, t, f->m);
    } else {
        memcpy(out, t, 32);
    }
}

static void fe_add(const field_params_t *f, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
    uint64_t t[4];
    uint64_t carry = add4(t, a, b);
    if (carry || ge4(t, f->m)) sub4(out, t, f->m); else memcpy(out, t, 32);
}

static void fe_sub(const field_params_t *f, const uint64_t a[4], const uint64_t b[4], uint64_t out[4]) {
    uint64_t t[4];
    if (sub4(t, a, b)) add4(out, t, f->m); else memcpy(out, t, 32);
}

static int fe_is_zero(const uint64_t a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }

/* a^(m-2) by square-and-multiply; 0 -> 0. */
static void fe_inv(const field_params_t *f, const uint64_t a[4], uint64_t out[4]) {
    uint64_t e[4], two[4] = { 2, 0, 0, 0 };
    sub4(e, f->m, two);
    uint64_t acc[4], base[4];
    memcpy(acc, f->r, 32);
    memcpy(base, a, 32);
    for (int i = 0; i < 256; ++i) {
        if ((e[i / 64] >> (i % 64)) & 1) mont_mul(f, acc, base, acc);
        mont_mul(f, base, base, base);
    }
    memcpy(out, acc, 32);
}

void oracle_fe_mul(int w, const ofe_t *a, const ofe_t *b, ofe_t *o) { mont_mul(&FIELDS[w], a->l, b->l, o->l); }
void oracle_fe_add(int w, const ofe_t *a, const ofe_t *b, ofe_t *o) { fe_add(&FIELDS[w], a->l, b->l, o->l); }
void oracle_fe_sub(int w, const ofe_t *a, const ofe_t *b, ofe_t *o) { fe_sub(&FIELDS[w], a->l, b->l, o->l); }
void oracle_fe_inv(int w, const ofe_t *a, ofe_t *o) { fe_inv(&FIELDS[w], a->l, o->l); }

void oracle_fe_from_canonical(int w, const uint64_t c[4], ofe_t *out) {
    mont_mul(&FIELDS[w], c, FIELDS[w].r2, out->l);
}

/* halo2_curves `to_repr` [ext]: one Montgomery reduction, little-endian bytes. */
void oracle_fe_to_canonical(int w, const ofe_t *a, uint64_t c[4]) {
    static const uint64_t one[4] = { 1, 0, 0, 0 };
    mont_mul(&FIELDS[w], a->l, one, c);
}

/* ------------------------------------------------------------------- curve */

#define FQ (&FIELDS[0])
#define FR (&FIELDS[1])

static void q_mul(const ofe_t *a, const ofe_t *b, ofe_t *o) { mont_mul(FQ, a->l, b->l, o->l); }
static void q_sqr(const ofe_t *a, ofe_t *o) { mont_mul(FQ, a->l, a->l, o->l); }
static void q_add(const ofe_t *a, const ofe_t *b, ofe_t *o) { fe_add(FQ, a->l, b->l, o->l); }
static void q_sub(const ofe_t *a, const ofe_t *b, ofe_t *o) { fe_sub(FQ, a->l, b->l, o->l); }
static void q_dbl(const ofe_t *a, ofe_t *o) { fe_add(FQ, a->l, a->l, o->l); }

static int affine_is_identity(const og1_affine_t *p) { return fe_is_zero(p->x.l) && fe_is_zero(p->y.l); }
static int jac_is_identity(const og1_jac_t *p) { return fe_is_zero(p->z.l); }

static void jac_set_identity(og1_jac_t *p) {
    memset(p, 0, sizeof(*p));
    memcpy(p->y.l, FQ->r, 32); /* (0 : 1 : 0) */
}

void oracle_g1_generator(og1_affine_t *out) {
    static const uint64_t one[4] = { 1, 0, 0, 0 }, two[4] = { 2, 0, 0, 0 };
    oracle_fe_from_canonical(0, one, &out->x);
    oracle_fe_from_canonical(0, two, &out->y);
}

int oracle_g1_is_on_curve(const og1_affine_t *p) {
    if (affine_is_identity(p)) return 1;
    static const uint64_t three[4] = { 3, 0, 0, 0 };
    ofe_t b, y2, x3;
    oracle_fe_from_canonical(0, three, &b);
    q_sqr(&p->y, &y2);
    q_sqr(&p->x, &x3);
    q_mul(&x3, &p->x, &x3);
    q_add(&x3, &b, &x3);
    return memcmp(&y2, &x3, 32) == 0;
}

void oracle_g1_from_affine(const og1_affine_t *a, og1_jac_t *out) {
    if (affine_is_identity(a)) { jac_set_identity(out); return; }
    out->x = a->x;
    out->y = a->y;
    memcpy(out->z.l, FQ->r, 32);
}

/* dbl-2009-l (a = 0): 2M + 5S. */
void oracle_g1_double(const og1_jac_t *p, og1_jac_t *out) {
    if (jac_is_identity(p)) { jac_set_identity(out); return; }
    ofe_t a, b, c, d, e, f, t, z3;
    q_sqr(&p->x, &a);
    q_sqr(&p->y, &b);
    q_sqr(&b, &c);
    q_add(&p->x, &b, &d);
    q_sqr(&d, &d);
    q_sub(&d, &a, &d);
    q_sub(&d, &c, &d);
    q_dbl(&d, &d);
    q_dbl(&a, &e);
    q_add(&e, &a, &e);
    q_sqr(&e, &f);
    q_mul(&p->y, &p->z, &z3);
    q_dbl(&z3, &z3);
    q_dbl(&d, &t);
    q_sub(&f, &t, &out->x);
    q_dbl(&c, &c); q_dbl(&c, &c); q_dbl(&c, &c);
    q_sub(&d, &out->x, &t);
    q_mul(&e, &t, &t);
    q_sub(&t, &c, &out->y);
    out->z = z3;
}

/* add-2007-bl with the exceptional cases (identity operands, P+P, P+(-P)). */
void oracle_g1_add(const og1_jac_t *p, const og1_jac_t *q, og1_jac_t *out) {
    if (jac_is_identity(p)) { *out = *q; return; }
    if (jac_is_identity(q)) { *out = *p; return; }
    ofe_t z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
    q_sqr(&p->z, &z1z1);
    q_sqr(&q->z, &z2z2);
    q_mul(&p->x, &z2z2, &u1);
    q_mul(&q->x, &z1z1, &u2);
    q_mul(&p->y, &q->z, &s1); q_mul(&s1, &z2z2, &s1);
    q_mul(&q->y, &p->z, &s2); q_mul(&s2, &z1z1, &s2);
    if (memcmp(&u1, &u2, 32) == 0) {
        if (memcmp(&s1, &s2, 32) == 0) { oracle_g1_double(p, out); } else { jac_set_identity(out); }
        return;
    }
    q_sub(&u2, &u1, &h);
    q_dbl(&h, &i); q_sqr(&i, &i);
    q_mul(&h, &i, &j);
    q_sub(&s2, &s1, &rr); q_dbl(&rr, &rr);
    q_mul(&u1, &i, &v);
    og1_jac_t o;
    q_sqr(&rr, &o.x); q_sub(&o.x, &j, &o.x); q_dbl(&v, &t); q_sub(&o.x, &t, &o.x);
    q_sub(&v, &o.x, &t); q_mul(&rr, &t, &t);
    q_mul(&s1, &j, &s1); q_dbl(&s1, &s1);
    q_sub(&t, &s1, &o.y);
    q_add(&p->z, &q->z, &o.z); q_sqr(&o.z, &o.z); q_sub(&o.z, &z1z1, &o.z); q_sub(&o.z, &z2z2, &o.z);
    q_mul(&o.z, &h, &o.z);
    *out = o;
}

/* madd-2007-bl with the exceptional cases. */
void oracle_g1_add_mixed(const og1_jac_t *p, const og1_affine_t *q, og1_jac_t *out) {
    if (affine_is_identity(q)) { *out = *p; return; }
    if (jac_is_identity(p)) { oracle_g1_from_affine(q, out); return; }
    ofe_t z1z1, u2, s2, h, hh, i, j, rr, v, t;
    q_sqr(&p->z, &z1z1);
    q_mul(&q->x, &z1z1, &u2);
    q_mul(&q->y, &p->z, &s2); q_mul(&s2, &z1z1, &s2);
    if (memcmp(&p->x, &u2, 32) == 0) {
        if (memcmp(&p->y, &s2, 32) == 0) { oracle_g1_double(p, out); } else { jac_set_identity(out); }
        return;
    }
    q_sub(&u2, &p->x, &h);
    q_sqr(&h, &hh);
    q_dbl(&hh, &i); q_dbl(&i, &i);
    q_mul(&h, &i, &j);
    q_sub(&s2, &p->y, &rr); q_dbl(&rr, &rr);
    q_mul(&p->x, &i, &v);
    og1_jac_t o;
    q_sqr(&rr, &o.x); q_sub(&o.x, &j, &o.x); q_dbl(&v, &t); q_sub(&o.x, &t, &o.x);
    q_sub(&v, &o.x, &t); q_mul(&rr, &t, &t);
    ofe_t yj; q_mul(&p->y, &j, &yj); q_dbl(&yj, &yj);
    q_sub(&t, &yj, &o.y);
    q_add(&p->z, &h, &o.z); q_sqr(&o.z, &o.z); q_sub(&o.z, &z1z1, &o.z); q_sub(&o.z, &hh, &o.z);
    *out = o;
}

void oracle_g1_to_affine(const og1_jac_t *p, og1_affine_t *out) {
    if (jac_is_identity(p)) { memset(out, 0, sizeof(*out)); return; }
    ofe_t zi, zi2, zi3;
    fe_inv(FQ, p->z.l, zi.l);
    q_sqr(&zi, &zi2);
    q_mul(&zi2, &zi, &zi3);
    q_mul(&p->x, &zi2, &out->x);
    q_mul(&p->y, &zi3, &out->y);
}

void oracle_g1_scalar_mul(const og1_affine_t *base, const uint64_t k[4], og1_jac_t *out) {
    og1_jac_t acc;
    jac_set_identity(&acc);
    for (int i = 255; i >= 0; --i) {
        oracle_g1_double(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) oracle_g1_add_mixed(&acc, base, &acc);
    }
    *out = acc;
}

int oracle_g1_transcript_bytes(const og1_affine_t *p, uint8_t out[64]) {
    if (affine_is_identity(p)) return -1;
    const ofe_t *coord[2] = { &p->x, &p->y };
    for (int c = 0; c < 2; ++c) {
        uint64_t canon[4];
        oracle_fe_to_canonical(0, coord[c], canon);
        for (int i = 0; i < 32; ++i) {
            /* to_repr() is little-endian; the transcript reverses it (transcript.rs:221-222). */
            out[32 * c + i] = (uint8_t)(canon[(31 - i) / 8] >> (8 * ((31 - i) % 8)));
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- hot path */

/* msm.rs:8-14 */
size_t oracle_window_size(size_t num_scalars) {
    if (num_scalars < 32) return 3;
    return (size_t)floor(log((double)num_scalars));
}

/* msm.rs:33-48 */
size_t oracle_windowed_scalar(size_t window_size, size_t window_mask, size_t idx, const uint8_t repr[32]) {
    size_t skip_bits = idx * window_size;
    size_t skip_bytes = skip_bits / 8;
    uint8_t value[8] = { 0 };
    for (size_t k = 0; k < 8 && skip_bytes + k < 32; ++k) value[k] = repr[skip_bytes + k];
    uint64_t v = 0;
    for (int k = 7; k >= 0; --k) v = (v << 8) | value[k];
    return (size_t)((v >> (skip_bits - skip_bytes * 8)) & window_mask);
}

/* msm.rs:122-151 — enum CurveAcc { Empty, Affine(C), Projective(C::Curve) } */
typedef struct {
    int tag; /* 0 Empty, 1 Affine, 2 Projective */
    og1_affine_t affine;
    og1_jac_t projective;
} curve_acc_t;

/* msm.rs:130-139 */
static void curve_acc_add_assign(curve_acc_t *acc, const og1_affine_t *rhs) {
    switch (acc->tag) {
    case 0:
        acc->tag = 1;
        acc->affine = *rhs;
        break;
    case 1: {
        og1_jac_t lhs;
        oracle_g1_from_affine(&acc->affine, &lhs);       /* affine + affine -> projective */
        oracle_g1_add_mixed(&lhs, rhs, &acc->projective);
        acc->tag = 2;
        break;
    }
    default:
        oracle_g1_add_mixed(&acc->projective, rhs, &acc->projective);
    }
}

/* msm.rs:141-150 */
static void curve_acc_add(const curve_acc_t *acc, og1_jac_t *rhs) {
    switch (acc->tag) {
    case 0: break;
    case 1: oracle_g1_add_mixed(rhs, &acc->affine, rhs); break;
    default: oracle_g1_add(&acc->projective, rhs, rhs);
    }
}

/* msm.rs:117-181 */
void oracle_variable_base_msm_serial(const ofe_t *scalars, const og1_affine_t *bases, size_t n,
                                     og1_jac_t *result) {
    if (n == 0) return;
    uint8_t *reprs = (uint8_t *)malloc(32 * n);                          /* :153 */
    for (size_t i = 0; i < n; ++i) {
        uint64_t canon[4];
        oracle_fe_to_canonical(1, &scalars[i], canon);
        memcpy(reprs + 32 * i, canon, 32);                               /* x86: little-endian */
    }
    const size_t num_bits = 8 * 32;                                      /* :154-155 */
    const size_t window_size = oracle_window_size(n);                    /* :157 */
    const size_t num_buckets = ((size_t)1 << window_size) - 1;           /* :158 */
    const size_t num_windows = (num_bits + window_size - 1) / window_size; /* :160 */
    curve_acc_t *buckets = (curve_acc_t *)malloc(sizeof(curve_acc_t) * num_buckets);

    for (size_t idx = num_windows; idx-- > 0;) {                         /* :161 */
        for (size_t k = 0; k < window_size; ++k) oracle_g1_double(result, result); /* :162-164 */
        for (size_t b = 0; b < num_buckets; ++b) buckets[b].tag = 0;     /* :166 */
        for (size_t i = 0; i < n; ++i) {                                 /* :168-173 */
            size_t s = oracle_windowed_scalar(window_size, num_buckets, idx, reprs + 32 * i);
            if (s != 0) curve_acc_add_assign(&buckets[s - 1], &bases[i]);
        }
        og1_jac_t running_sum;                                           /* :175-179 */
        jac_set_identity(&running_sum);
        for (size_t b = num_buckets; b-- > 0;) {
            curve_acc_add(&buckets[b], &running_sum);
            oracle_g1_add(result, &running_sum, result);
        }
    }
    free(buckets);
    free(reprs);
}

typedef struct {
    const ofe_t *scalars;
    const og1_affine_t *bases;
    size_t n;
    og1_jac_t result;
} msm_task_t;

static void *msm_task_run(void *arg) {
    msm_task_t *t = (msm_task_t *)arg;
    oracle_variable_base_msm_serial(t->scalars, t->bases, t->n, &t->result);
    return NULL;
}

/* msm.rs:84-115, with parallel.rs:9-25 restated as one pthread per chunk. */
void oracle_variable_base_msm(const ofe_t *scalars, const og1_affine_t *bases, size_t n,
                              int num_threads, og1_jac_t *out) {
    jac_set_identity(out);
    if (n == 0) return;
    if (num_threads < 1) num_threads = 1;
    if (n <= (size_t)num_threads) {                                      /* :95-99 */
        oracle_variable_base_msm_serial(scalars, bases, n, out);
        return;
    }
    size_t chunk_size = (n + num_threads - 1) / num_threads;             /* :101 */
    size_t num_chunks = (n + chunk_size - 1) / chunk_size;
    msm_task_t *tasks = (msm_task_t *)calloc(num_chunks, sizeof(msm_task_t));
    pthread_t *threads = (pthread_t *)calloc(num_chunks, sizeof(pthread_t));
    for (size_t c = 0; c < num_chunks; ++c) {                            /* :103-111 */
        size_t start = c * chunk_size;
        tasks[c].scalars = scalars + start;
        tasks[c].bases = bases + start;
        tasks[c].n = (start + chunk_size <= n) ? chunk_size : n - start;
        jac_set_identity(&tasks[c].result);
        pthread_create(&threads[c], NULL, msm_task_run, &tasks[c]);
    }
    for (size_t c = 0; c < num_chunks; ++c) {
        pthread_join(threads[c], NULL);
        oracle_g1_add(out, &tasks[c].result, out);                       /* :112-114 */
    }
    free(threads);
    free(tasks);
}

void oracle_msm_naive(const ofe_t *scalars, const og1_affine_t *bases, size_t n, og1_jac_t *out) {
    jac_set_identity(out);
    for (size_t i = 0; i < n; ++i) {
        uint64_t k[4];
        og1_jac_t t;
        oracle_fe_to_canonical(1, &scalars[i], k);
        oracle_g1_scalar_mul(&bases[i], k, &t);
        oracle_g1_add(out, &t, out);
    }
}

/* ------------------------------------------------- known-discrete-log inputs */

/* Normalise jac[0..n) to affine with one inversion (Montgomery's trick). */
static void batch_to_affine(const og1_jac_t *jac, size_t n, og1_affine_t *out) {
    if (n == 0) return;
    ofe_t *prefix = (ofe_t *)malloc(sizeof(ofe_t) * n);
    ofe_t acc;
    memcpy(acc.l, FQ->r, 32);
    for (size_t i = 0; i < n; ++i) {
        prefix[i] = acc;
        if (!jac_is_identity(&jac[i])) q_mul(&acc, &jac[i].z, &acc);
    }
    ofe_t inv;
    fe_inv(FQ, acc.l, inv.l);
    for (size_t i = n; i-- > 0;) {
        if (jac_is_identity(&jac[i])) { memset(&out[i], 0, sizeof(out[i])); continue; }
        ofe_t zi, zi2, zi3;
        q_mul(&inv, &prefix[i], &zi);
        q_mul(&inv, &jac[i].z, &inv);
        q_sqr(&zi, &zi2);
        q_mul(&zi2, &zi, &zi3);
        q_mul(&jac[i].x, &zi2, &out[i].x);
        q_mul(&jac[i].y, &zi3, &out[i].y);
    }
    free(prefix);
}

typedef struct {
    uint64_t a[4], d[4];
    size_t start, count;
    og1_affine_t step; /* d*G */
    og1_affine_t *out;
} dlog_task_t;

static void *dlog_task_run(void *arg) {
    dlog_task_t *t = (dlog_task_t *)arg;
    if (t->count == 0) return NULL;
    /* k0 = a + start*d (mod r) */
    ofe_t fa, fd, fs, k0m;
    uint64_t s[4] = { (uint64_t)t->start, 0, 0, 0 }, k0[4];
    oracle_fe_from_canonical(1, t->a, &fa);
    oracle_fe_from_canonical(1, t->d, &fd);
    oracle_fe_from_canonical(1, s, &fs);
    mont_mul(FR, fs.l, fd.l, k0m.l);
    fe_add(FR, k0m.l, fa.l, k0m.l);
    oracle_fe_to_canonical(1, &k0m, k0);
    og1_affine_t g;
    oracle_g1_generator(&g);
    const size_t BATCH = 1024;
    og1_jac_t *buf = (og1_jac_t *)malloc(sizeof(og1_jac_t) * BATCH);
    og1_jac_t cur;
    oracle_g1_scalar_mul(&g, k0, &cur);
    size_t done = 0;
    while (done < t->count) {
        size_t m = t->count - done < BATCH ? t->count - done : BATCH;
        for (size_t i = 0; i < m; ++i) {
            buf[i] = cur;
            oracle_g1_add_mixed(&cur, &t->step, &cur);
        }
        batch_to_affine(buf, m, t->out + done);
        done += m;
    }
    free(buf);
    return NULL;
}

void oracle_known_dlog_bases(const uint64_t a[4], const uint64_t d[4], size_t n, int num_threads,
                             og1_affine_t *out) {
    if (n == 0) return;
    if (num_threads < 1) num_threads = 1;
    if ((size_t)num_threads > n) num_threads = (int)n;
    og1_affine_t g, step;
    og1_jac_t dj;
    oracle_g1_generator(&g);
    oracle_g1_scalar_mul(&g, d, &dj);
    oracle_g1_to_affine(&dj, &step);
    dlog_task_t *tasks = (dlog_task_t *)calloc(num_threads, sizeof(dlog_task_t));
    pthread_t *threads = (pthread_t *)calloc(num_threads, sizeof(pthread_t));
    size_t chunk = (n + num_threads - 1) / num_threads;
    for (int t = 0; t < num_threads; ++t) {
        size_t start = (size_t)t * chunk;
        memcpy(tasks[t].a, a, 32);
        memcpy(tasks[t].d, d, 32);
        tasks[t].start = start;
        tasks[t].count = start >= n ? 0 : (start + chunk <= n ? chunk : n - start);
        tasks[t].step = step;
        tasks[t].out = out + start;
        pthread_create(&threads[t], NULL, dlog_task_run, &tasks[t]);
    }
    for (int t = 0; t < num_threads; ++t) pthread_join(threads[t], NULL);
    free(threads);
    free(tasks);
}

void oracle_known_dlog_answer(const uint64_t a[4], const uint64_t d[4], const ofe_t *scalars, size_t n,
                              og1_affine_t *out) {
    ofe_t fa, fd, sum, wsum, idx, one, t;
    static const uint64_t c1[4] = { 1, 0, 0, 0 };
    oracle_fe_from_canonical(1, a, &fa);
    oracle_fe_from_canonical(1, d, &fd);
    oracle_fe_from_canonical(1, c1, &one);
    memset(&sum, 0, sizeof(sum));
    memset(&wsum, 0, sizeof(wsum));
    memset(&idx, 0, sizeof(idx));
    for (size_t i = 0; i < n; ++i) {
        fe_add(FR, sum.l, scalars[i].l, sum.l);
        mont_mul(FR, idx.l, scalars[i].l, t.l);
        fe_add(FR, wsum.l, t.l, wsum.l);
        fe_add(FR, idx.l, one.l, idx.l);
    }
    mont_mul(FR, fa.l, sum.l, sum.l);
    mont_mul(FR, fd.l, wsum.l, wsum.l);
    fe_add(FR, sum.l, wsum.l, sum.l);
    uint64_t k[4];
    oracle_fe_to_canonical(1, &sum, k);
    og1_affine_t g;
    og1_jac_t rj;
    oracle_g1_generator(&g);
    oracle_g1_scalar_mul(&g, k, &rj);
    oracle_g1_to_affine(&rj, out);
}

/* ------------------------------------------- callers either side of the MSM */

/* pcs/multilinear.rs:72-107.  quotients[] holds 2^num_vars entries; q_i (2^i values,
 * committed against eqs[i] at kzg.rs:291-293) sits at offset 2^i, entry 0 is unused. */
void oracle_quotients(const ofe_t *evals, const ofe_t *point, size_t num_vars, ofe_t *quotients, ofe_t *eval_out) {
    const size_t n = (size_t)1 << num_vars;
    ofe_t *remainder = (ofe_t *)malloc(sizeof(ofe_t) * n);
    memcpy(remainder, evals, sizeof(ofe_t) * n);
    memset(&quotients[0], 0, sizeof(ofe_t));
    for (size_t i = num_vars; i-- > 0;) {            /* .zip(0..num_vars).rev()           (:80-83) */
        const size_t half = (size_t)1 << i;
        ofe_t *q = quotients + half;
        for (size_t j = 0; j < half; ++j) {
            fe_sub(FR, remainder[half + j].l, remainder[j].l, q[j].l);      /* *q = *r_hi - r_lo   (:90-91) */
            ofe_t t;
            mont_mul(FR, q[j].l, point[i].l, t.l);                          /* (r_hi - r_lo) * x_i (:94-95) */
            fe_add(FR, remainder[j].l, t.l, remainder[j].l);
        }
    }
    *eval_out = remainder[0];
    free(remainder);
}

/* pcs/multilinear.rs:203-213 ("g_prime"): sum_i coeffs[i] * polys[i], element-wise. */
void oracle_fr_linear_combination(const ofe_t *const *polys, const ofe_t *coeffs, size_t count, size_t n, ofe_t *out) {
    for (size_t j = 0; j < n; ++j) {
        ofe_t acc, t;
        memset(&acc, 0, sizeof(acc));
        for (size_t i = 0; i < count; ++i) {
            mont_mul(FR, coeffs[i].l, polys[i][j].l, t.l);
            fe_add(FR, acc.l, t.l, acc.l);
        }
        out[j] = acc;
    }
}

/* Element-wise Fr vector operations on all host threads (bench input generation and the host-side tables of the
 * compiled sum-check expression): op 0: a + b, 1: a - b, 2: a * b;
 * oracle_fr_affine: out[j] = constant + id_coeff * j + sum_i coeffs[i] * polys[i][rows_i ? rows_i[j] : j] — the explicit
 * form of a linear factor of the zero-check expression (backend/hyperplonk/preprocessor.rs:153-165) with its identity
 * polynomial (piop/sum_check/classic.rs:92) and rotated queries (classic.rs:105-125). */
typedef struct { int op; const ofe_t *a, *b; size_t n; ofe_t *out; } vec_task_t;
static void *vec_task_run(void *arg) {
    vec_task_t *t = (vec_task_t *)arg;
    for (size_t j = 0; j < t->n; ++j) {
        if (t->op == 0) fe_add(FR, t->a[j].l, t->b[j].l, t->out[j].l);
        else if (t->op == 1) fe_sub(FR, t->a[j].l, t->b[j].l, t->out[j].l);
        else mont_mul(FR, t->a[j].l, t->b[j].l, t->out[j].l);
    }
    return NULL;
}
void oracle_fr_vec_op(int op, const ofe_t *a, const ofe_t *b, size_t n, int num_threads, ofe_t *out) {
    if (num_threads < 1) num_threads = 1;
    if ((size_t)num_threads > n) num_threads = n ? (int)n : 1;
    const size_t per = (n + num_threads - 1) / num_threads;
    vec_task_t *tasks = (vec_task_t *)calloc(num_threads, sizeof(vec_task_t));
    pthread_t *threads = (pthread_t *)calloc(num_threads, sizeof(pthread_t));
    for (int t = 0; t < num_threads; ++t) {
        const size_t first = (size_t)t * per;
        const size_t count = first >= n ? 0 : (first + per <= n ? per : n - first);
        vec_task_t k = {op, a + first, b + first, count, out + first};
        tasks[t] = k;
        pthread_create(&threads[t], NULL, vec_task_run, &tasks[t]);
    }
    for (int t = 0; t < num_threads; ++t) pthread_join(threads[t], NULL);
    free(tasks); free(threads);
}
typedef struct {
    const ofe_t *const *polys; const uint32_t *const *rows; const ofe_t *coeffs; size_t count;
    const ofe_t *constant, *id_coeff; size_t first, n; ofe_t *out;
} aff_task_t;
static void *aff_task_run(void *arg) {
    aff_task_t *t = (aff_task_t *)arg;
    for (size_t j = t->first; j < t->first + t->n; ++j) {
        ofe_t acc, tmp;
        memset(&acc, 0, sizeof(acc));
        if (t->constant) acc = *t->constant;
        if (t->id_coeff) {
            uint64_t c[4] = {(uint64_t)j, 0, 0, 0};
            oracle_fe_from_canonical(1, c, &tmp);
            mont_mul(FR, tmp.l, t->id_coeff->l, tmp.l);
            fe_add(FR, acc.l, tmp.l, acc.l);
        }
        for (size_t i = 0; i < t->count; ++i) {
            const size_t row = (t->rows && t->rows[i]) ? t->rows[i][j] : j;
            mont_mul(FR, t->coeffs[i].l, t->polys[i][row].l, tmp.l);
            fe_add(FR, acc.l, tmp.l, acc.l);
        }
        t->out[j] = acc;
    }
    return NULL;
}
void oracle_fr_affine(const ofe_t *const *polys, const uint32_t *const *rows, const ofe_t *coeffs, size_t count, const ofe_t *constant,
                      const ofe_t *id_coeff, size_t n, int num_threads, ofe_t *out) {
    if (num_threads < 1) num_threads = 1;
    if ((size_t)num_threads > n) num_threads = n ? (int)n : 1;
    const size_t per = (n + num_threads - 1) / num_threads;
    aff_task_t *tasks = (aff_task_t *)calloc(num_threads, sizeof(aff_task_t));
    pthread_t *threads = (pthread_t *)calloc(num_threads, sizeof(pthread_t));
    for (int t = 0; t < num_threads; ++t) {
        const size_t first = (size_t)t * per;
        const size_t cnt = first >= n ? 0 : (first + per <= n ? per : n - first);
        aff_task_t k = {polys, rows, coeffs, count, constant, id_coeff, first, cnt, out};
        tasks[t] = k;
        pthread_create(&threads[t], NULL, aff_task_run, &tasks[t]);
    }
    for (int t = 0; t < num_threads; ++t) pthread_join(threads[t], NULL);
    free(tasks); free(threads);
}

/* pcs/multilinear/kzg.rs:174-193: eqs[0] = [1]; eqs[k+1] = [e - s_k*e for e in eqs[k]] ++ [s_k*e ...].
 * out holds 2^(num_vars+1) - 1 scalars, slice k at offset 2^k - 1 (the flat_map order of :199). */
void oracle_kzg_eq_scalars(const ofe_t *ss, size_t num_vars, ofe_t *out) {
    memcpy(out[0].l, FR->r, 32);
    for (size_t k = 0; k < num_vars; ++k) {
        const size_t len = (size_t)1 << k;
        const ofe_t *last = out + (len - 1);
        ofe_t *lo = out + (2 * len - 1), *hi = lo + len;
        for (size_t j = 0; j < len; ++j) {
            mont_mul(FR, ss[k].l, last[j].l, hi[j].l);        /* *eval_hi = *s_i * last_eval        (:186) */
            fe_sub(FR, last[j].l, hi[j].l, lo[j].l);          /* *eval_lo = *last_eval - eval_hi    (:190) */
        }
    }
}

/* msm.rs:16-31 window_table + :50-81 fixed_base_msm, then batch_normalize (kzg.rs:204-207). */
typedef struct {
    size_t window_size, num_windows, start, count;
    const og1_affine_t *table; /* num_windows rows of (2^window_size - 1) points */
    const ofe_t *scalars;
    og1_affine_t *out;
} fixed_task_t;

static void *fixed_task_run(void *arg) {
    fixed_task_t *t = (fixed_task_t *)arg;
    const size_t row = ((size_t)1 << t->window_size) - 1, mask = row;
    const size_t BATCH = 1024;
    og1_jac_t *buf = (og1_jac_t *)malloc(sizeof(og1_jac_t) * BATCH);
    for (size_t done = 0; done < t->count;) {
        const size_t m = t->count - done < BATCH ? t->count - done : BATCH;
        for (size_t i = 0; i < m; ++i) {
            uint64_t c[4];
            uint8_t repr[32];
            oracle_fe_to_canonical(1, &t->scalars[t->start + done + i], c);
            memcpy(repr, c, 32);
            og1_jac_t acc;
            jac_set_identity(&acc);
            for (size_t w = 0; w < t->num_windows; ++w) {                     /* windowed_scalar_mul (:50-65) */
                const size_t d = oracle_windowed_scalar(t->window_size, mask, w, repr);
                if (d > 0) oracle_g1_add_mixed(&acc, &t->table[w * row + d - 1], &acc);
            }
            buf[i] = acc;
        }
        batch_to_affine(buf, m, t->out + t->start + done);
        done += m;
    }
    free(buf);
    return NULL;
}

void oracle_fixed_base_msm(const og1_affine_t *base, size_t window_size, const ofe_t *scalars, size_t n, int num_threads,
                           og1_affine_t *out) {
    if (n == 0) return;
    if (num_threads < 1) num_threads = 1;
    if ((size_t)num_threads > n) num_threads = (int)n;
    const size_t scalar_size = 254;                                           /* field_size::<Fr>() (arithmetic.rs:202-205) */
    const size_t num_windows = (scalar_size + window_size - 1) / window_size;
    const size_t row = ((size_t)1 << window_size) - 1;
    og1_affine_t *table = (og1_affine_t *)malloc(sizeof(og1_affine_t) * num_windows * row);
    og1_jac_t *tmp = (og1_jac_t *)malloc(sizeof(og1_jac_t) * row);
    og1_jac_t offset;
    oracle_g1_from_affine(base, &offset);
    for (size_t w = 0; w < num_windows; ++w) {                                /* window_table (:16-31) */
        og1_affine_t off_aff;
        oracle_g1_to_affine(&offset, &off_aff);
        og1_jac_t acc = offset;
        for (size_t v = 0; v < row; ++v) {
            tmp[v] = acc;
            oracle_g1_add_mixed(&acc, &off_aff, &acc);
        }
        batch_to_affine(tmp, row, table + w * row);
        for (size_t b = 0; b < window_size; ++b) oracle_g1_double(&offset, &offset);
    }
    free(tmp);
    fixed_task_t *tasks = (fixed_task_t *)calloc(num_threads, sizeof(fixed_task_t));
    pthread_t *threads = (pthread_t *)calloc(num_threads, sizeof(pthread_t));
    const size_t chunk = (n + num_threads - 1) / num_threads;
    for (int t = 0; t < num_threads; ++t) {
        const size_t start = (size_t)t * chunk;
        tasks[t].window_size = window_size;
        tasks[t].num_windows = num_windows;
        tasks[t].start = start;
        tasks[t].count = start >= n ? 0 : (start + chunk <= n ? chunk : n - start);
        tasks[t].table = table;
        tasks[t].scalars = scalars;
        tasks[t].out = out;
        pthread_create(&threads[t], NULL, fixed_task_run, &tasks[t]);
    }
    for (int t = 0; t < num_threads; ++t) pthread_join(threads[t], NULL);
    free(threads);
    free(tasks);
    free(table);
}

/* ----------------------------------------------------------------- sum check */

/* piop/sum_check/classic/eval.rs:101-131 on explicit tables: for every pair b the tables are read at
 * (2b, 2b+1) (eval.rs:236-243), the point X = 1 is e[2b+1] and every further point adds the step
 * e[2b+1] - e[2b] (eval.rs:268-300); out[x-1] = sum_b expr(X = x) for x = 1..degree.  expr =
 * sum_t coeffs[t] * prod_j polys[term_polys[j]] (j in [offsets[t], offsets[t+1])), times polys[common]
 * when common >= 0. */
void oracle_sumcheck_round(const ofe_t *const *polys, size_t num_polys, size_t size, const ofe_t *coeffs,
                           const uint32_t *offsets, const uint32_t *term_polys, size_t num_terms, int common,
                           size_t degree, ofe_t *out) {
    ofe_t *val = (ofe_t *)malloc(sizeof(ofe_t) * num_polys), *step = (ofe_t *)malloc(sizeof(ofe_t) * num_polys);
    memset(out, 0, sizeof(ofe_t) * degree);
    for (size_t b = 0; b < size; ++b) {
        for (size_t p = 0; p < num_polys; ++p) {
            val[p] = polys[p][2 * b + 1];
            fe_sub(FR, polys[p][2 * b + 1].l, polys[p][2 * b].l, step[p].l);
        }
        for (size_t x = 0; x < degree; ++x) {
            ofe_t total;
            memset(&total, 0, sizeof(total));
            for (size_t t = 0; t < num_terms; ++t) {
                ofe_t prod = coeffs[t];
                for (uint32_t j = offsets[t]; j < offsets[t + 1]; ++j) mont_mul(FR, prod.l, val[term_polys[j]].l, prod.l);
                fe_add(FR, total.l, prod.l, total.l);
            }
            if (common >= 0) mont_mul(FR, total.l, val[common].l, total.l);
            fe_add(FR, out[x].l, total.l, out[x].l);
            for (size_t p = 0; p < num_polys; ++p) fe_add(FR, val[p].l, step[p].l, val[p].l);
        }
    }
    free(val);
    free(step);
}

/* MultilinearPolynomial::fix_var (poly/multilinear.rs:179-189, merge_into :599-618):
 * out[b] = (e[2b+1] - e[2b]) * x + e[2b] for b < n/2. */
void oracle_fix_var(const ofe_t *evals, size_t n, const ofe_t *x, ofe_t *out) {
    for (size_t b = 0; b < n / 2; ++b) {
        ofe_t d;
        fe_sub(FR, evals[2 * b + 1].l, evals[2 * b].l, d.l);
        mont_mul(FR, d.l, x->l, d.l);
        fe_add(FR, d.l, evals[2 * b].l, out[b].l);
    }
}

/* Threaded forms of the two sum-check loops for the full-size checks in bench.py: the pair range is cut into
 * contiguous chunks, one pthread each (the way parallelize splits loops, util/parallel.rs:27-56); field addition is
 * exact, so the per-chunk partial sums add up to the same round message bit for bit. */
typedef struct {
    const ofe_t *const *polys; size_t num_polys, first, count; const ofe_t *coeffs; const uint32_t *offsets, *term_polys;
    size_t num_terms; int common; size_t degree; ofe_t *out;
} sc_task_t;
static void *sc_task_run(void *arg) {
    sc_task_t *t = (sc_task_t *)arg;
    const ofe_t **shifted = (const ofe_t **)malloc(sizeof(ofe_t *) * t->num_polys);
    for (size_t p = 0; p < t->num_polys; ++p) shifted[p] = t->polys[p] + 2 * t->first;
    oracle_sumcheck_round(shifted, t->num_polys, t->count, t->coeffs, t->offsets, t->term_polys, t->num_terms, t->common, t->degree, t->out);
    free(shifted);
    return NULL;
}
void oracle_sumcheck_round_mt(const ofe_t *const *polys, size_t num_polys, size_t size, const ofe_t *coeffs,
                              const uint32_t *offsets, const uint32_t *term_polys, size_t num_terms, int common,
                              size_t degree, int num_threads, ofe_t *out) {
    if (num_threads < 1) num_threads = 1;
    if ((size_t)num_threads > size) num_threads = size ? (int)size : 1;
    const size_t per = (size + num_threads - 1) / num_threads;
    sc_task_t *tasks = (sc_task_t *)calloc(num_threads, sizeof(sc_task_t));
    pthread_t *threads = (pthread_t *)calloc(num_threads, sizeof(pthread_t));
    ofe_t *partial = (ofe_t *)calloc((size_t)num_threads * degree, sizeof(ofe_t));
    for (int t = 0; t < num_threads; ++t) {
        const size_t first = (size_t)t * per;
        const size_t count = first >= size ? 0 : (first + per <= size ? per : size - first);
        sc_task_t k = {polys, num_polys, first, count, coeffs, offsets, term_polys, num_terms, common, degree, partial + (size_t)t * degree};
        tasks[t] = k;
        pthread_create(&threads[t], NULL, sc_task_run, &tasks[t]);
    }
    memset(out, 0, sizeof(ofe_t) * degree);
    for (int t = 0; t < num_threads; ++t) {
        pthread_join(threads[t], NULL);
        for (size_t x = 0; x < degree; ++x) fe_add(FR, out[x].l, partial[(size_t)t * degree + x].l, out[x].l);
    }
    free(tasks); free(threads); free(partial);
}
typedef struct { const ofe_t *evals; size_t n; const ofe_t *x; ofe_t *out; } fv_task_t;
static void *fv_task_run(void *arg) {
    fv_task_t *t = (fv_task_t *)arg;
    oracle_fix_var(t->evals, t->n, t->x, t->out);
    return NULL;
}
void oracle_fix_var_mt(const ofe_t *evals, size_t n, const ofe_t *x, int num_threads, ofe_t *out) {
    const size_t half = n / 2;
    if (num_threads < 1) num_threads = 1;
    if ((size_t)num_threads > half) num_threads = half ? (int)half : 1;
    const size_t per = (half + num_threads - 1) / num_threads;
    fv_task_t *tasks = (fv_task_t *)calloc(num_threads, sizeof(fv_task_t));
    pthread_t *threads = (pthread_t *)calloc(num_threads, sizeof(pthread_t));
    for (int t = 0; t < num_threads; ++t) {
        const size_t first = (size_t)t * per;
        const size_t count = first >= half ? 0 : (first + per <= half ? per : half - first);
        fv_task_t k = {evals + 2 * first, 2 * count, x, out + first};
        tasks[t] = k;
        pthread_create(&threads[t], NULL, fv_task_run, &tasks[t]);
    }
    for (int t = 0; t < num_threads; ++t) pthread_join(threads[t], NULL);
    free(tasks); free(threads);
}

/* ---------------------------------------------------------------- transcript */

/* Keccak-f[1600] and Keccak256 (original padding 0x01, rate 136): the hash behind Keccak256Transcript
 * (util/transcript.rs:100-131, util/hash.rs:5-8; the reference takes it from the sha3 crate [ext]).  Restated from
 * the published Keccak specification; pinned in the tests by the public digests of "" and "abc" and against the
 * independent Python implementation in plonkish_b200/transcript.py. */
static uint64_t rol64(uint64_t v, unsigned n) { return n ? (v << n) | (v >> (64 - n)) : v; }

static void keccak_f1600(uint64_t a[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000AULL,
        0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800AULL, 0x800000008000000AULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    /* rho offsets walked along the pi cycle starting at lane (1, 0) */
    static const unsigned RHO[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const unsigned PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; ++round) {
        uint64_t c[5];
        for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; ++x) {
            const uint64_t d = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
            for (int y = 0; y < 25; y += 5) a[y + x] ^= d;
        }
        uint64_t cur = a[1];
        for (int i = 0; i < 24; ++i) {
            const uint64_t next = a[PI[i]];
            a[PI[i]] = rol64(cur, RHO[i]);
            cur = next;
        }
        for (int y = 0; y < 25; y += 5) {
            uint64_t row[5];
            for (int x = 0; x < 5; ++x) row[x] = a[y + x];
            for (int x = 0; x < 5; ++x) a[y + x] = row[x] ^ (~row[(x + 1) % 5] & row[(x + 2) % 5]);
        }
        a[0] ^= RC[round];
    }
}

void oracle_keccak256(const uint8_t *data, size_t len, uint8_t out[32]) {
    uint64_t st[25];
    uint8_t block[136];
    memset(st, 0, sizeof(st));
    while (1) {
        const size_t take = len < 136 ? len : 136;
        memset(block, 0, sizeof(block));
        memcpy(block, data, take);
        const int last = take < 136;
        if (last) {
            block[take] ^= 0x01;
            block[135] ^= 0x80;
        }
        for (int i = 0; i < 17; ++i) {
            uint64_t lane = 0;
            for (int b = 7; b >= 0; --b) lane = (lane << 8) | block[8 * i + b];
            st[i] ^= lane;
        }
        keccak_f1600(st);
        if (last) break;
        data += take;
        len -= take;
    }
    for (int i = 0; i < 4; ++i)
        for (int b = 0; b < 8; ++b) out[8 * i + b] = (uint8_t)(st[i] >> (8 * b));
}

/* Array forms of the two conversions (the element-wise ctypes loop is too slow for the 2^24-element bench inputs). */
void oracle_fe_from_canonical_n(int which, const uint64_t *in, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; ++i) oracle_fe_from_canonical(which, in + 4 * i, (ofe_t *)(out + 4 * i));
}
void oracle_fe_to_canonical_n(int which, const uint64_t *in, size_t n, uint64_t *out) {
    for (size_t i = 0; i < n; ++i) oracle_fe_to_canonical(which, (const ofe_t *)(in + 4 * i), out + 4 * i);
}

/* UnivariatePolynomial::div_rem (poly/univariate.rs:144-168) for the divisor (X - z) of UnivariateKzg::open
 * (pcs/univariate/kzg.rs:281-282): the reference's long division, specialised to divisor coefficients [-z, 1]:
 * every step takes the leading remainder coefficient as the next quotient coefficient and subtracts
 * quotient_coeff * divisor shifted into place.  q receives n - 1 coefficients, rem the remainder (the value at z). */
void oracle_fr_div_linear(const ofe_t *coeffs, size_t n, const ofe_t *z, ofe_t *q, ofe_t *rem) {
    if (n == 0) { memset(rem, 0, sizeof(*rem)); return; }
    ofe_t *r = (ofe_t *)malloc(sizeof(ofe_t) * n);
    memcpy(r, coeffs, sizeof(ofe_t) * n);
    for (size_t degree = n - 1; degree-- > 0;) {   /* remainder.degree() - 1 = degree */
        const ofe_t qc = r[degree + 1];           /* divisor_leading_inv = 1 */
        ofe_t t;
        q[degree] = qc;
        mont_mul(FR, qc.l, z->l, t.l);            /* remainder[degree] -= qc * (-z) */
        fe_add(FR, r[degree].l, t.l, r[degree].l);
    }
    *rem = r[0];
    free(r);
}

/* ---- permutation_z_polys (backend/hyperplonk/prover.rs:252-345) with BooleanHypercube (util/arithmetic/bh.rs) -------- */
static const uint32_t BH_PRIMITIVES[32] = {1u, 3u, 7u, 11u, 19u, 37u, 67u, 131u, 285u, 529u, 1033u, 2053u, 4179u, 8219u, 16427u, 32771u, 65581u, 131081u,
                                           262183u, 524327u, 1048585u, 2097157u, 4194307u, 8388641u, 16777243u, 33554441u, 67108935u, 134217767u,
                                           268435465u, 536870917u, 1073741907u, 2147483657u}; /* bh.rs:5-38 */
static size_t bh_next(size_t b, size_t num_vars, size_t primitive) { /* bh.rs:141-146 */
    b <<= 1;
    b ^= (b >> num_vars) * primitive;
    return b;
}
/* BooleanHypercube::iter (bh.rs:118-125): 0, then 1 and its successors; out receives 2^num_vars entries. */
void oracle_bh_iter(size_t num_vars, uint32_t *out) {
    const size_t n = (size_t)1 << num_vars;
    size_t b = 1;
    out[0] = 0;
    for (size_t i = 1; i < n; ++i) {
        out[i] = (uint32_t)b;
        b = bh_next(b, num_vars, BH_PRIMITIVES[num_vars]);
    }
}
/* One chunk's products for rows [lo, hi) (prover.rs:267-299; `parallelize` cuts the rows the same way). */
typedef struct {
    const ofe_t *const *values, *const *sigmas; size_t first, last, num_vars, lo, hi; const ofe_t *beta, *gamma; ofe_t *product;
} perm_task_t;
static void *perm_task_run(void *arg) {
    perm_task_t *t = (perm_task_t *)arg;
    ofe_t *product = t->product;
    uint64_t one_c[4] = {1, 0, 0, 0};
    ofe_t one;
    oracle_fe_from_canonical(1, one_c, &one);
    for (size_t b = t->lo; b < t->hi; ++b) product[b] = one;
    for (size_t i = t->first; i < t->last; ++i)
        for (size_t b = t->lo; b < t->hi; ++b) {                                      /* product *= beta * permutation + gamma + value */
            ofe_t v;
            mont_mul(FR, t->beta->l, t->sigmas[i][b].l, v.l);
            fe_add(FR, v.l, t->gamma->l, v.l);
            fe_add(FR, v.l, t->values[i][b].l, v.l);
            mont_mul(FR, product[b].l, v.l, product[b].l);
        }
    /* batch_invert (prover.rs:280-283): Montgomery's trick per strip — the same inverses as one by one */
    {
        enum { STRIP = 256 };
        ofe_t prefix[STRIP];
        for (size_t s0 = t->lo; s0 < t->hi; s0 += STRIP) {
            const size_t cnt = t->hi - s0 < STRIP ? t->hi - s0 : STRIP;
            ofe_t run = one, inv;
            for (size_t j = 0; j < cnt; ++j) { prefix[j] = run; mont_mul(FR, run.l, product[s0 + j].l, run.l); }
            fe_inv(FR, run.l, inv.l);
            for (size_t j = cnt; j-- > 0;) {
                ofe_t v;
                mont_mul(FR, inv.l, prefix[j].l, v.l);
                mont_mul(FR, inv.l, product[s0 + j].l, inv.l);
                product[s0 + j] = v;
            }
        }
    }
    for (size_t i = t->first; i < t->last; ++i)
        for (size_t b = t->lo; b < t->hi; ++b) {                                      /* product *= beta * id + gamma + value, id = (idx << num_vars) + b */
            uint64_t id_c[4] = {((uint64_t)i << t->num_vars) + b, 0, 0, 0};
            ofe_t id, v;
            oracle_fe_from_canonical(1, id_c, &id);
            mont_mul(FR, t->beta->l, id.l, v.l);
            fe_add(FR, v.l, t->gamma->l, v.l);
            fe_add(FR, v.l, t->values[i][b].l, v.l);
            mont_mul(FR, product[b].l, v.l, product[b].l);
        }
    return NULL;
}
/* values[i], sigmas[i]: 2^num_vars Montgomery Fr each, i < count; out: num_chunks polynomials of 2^num_vars, one after another. */
void oracle_permutation_z_polys_mt(size_t num_chunks, const ofe_t *const *values, const ofe_t *const *sigmas, size_t count, size_t num_vars,
                                   const ofe_t *beta, const ofe_t *gamma, int num_threads, ofe_t *out) {
    const size_t n = (size_t)1 << num_vars;
    const size_t chunk_size = (count + num_chunks - 1) / num_chunks;                /* prover.rs:263 */
    ofe_t *products = (ofe_t *)malloc(sizeof(ofe_t) * num_chunks * n);
    uint64_t one_c[4] = {1, 0, 0, 0};
    ofe_t one;
    oracle_fe_from_canonical(1, one_c, &one);
    if (num_threads < 1) num_threads = 1;
    if ((size_t)num_threads > n) num_threads = (int)n;
    for (size_t chunk = 0; chunk < num_chunks; ++chunk) {
        const size_t first = chunk * chunk_size, last = first + chunk_size < count ? first + chunk_size : count;
        const size_t per = (n + num_threads - 1) / num_threads;
        perm_task_t *tasks = (perm_task_t *)calloc(num_threads, sizeof(perm_task_t));
        pthread_t *threads = (pthread_t *)calloc(num_threads, sizeof(pthread_t));
        for (int t = 0; t < num_threads; ++t) {
            const size_t lo = (size_t)t * per < n ? (size_t)t * per : n, hi = lo + per < n ? lo + per : n;
            perm_task_t k = {values, sigmas, first, last, num_vars, lo, hi, beta, gamma, products + chunk * n};
            tasks[t] = k;
            pthread_create(&threads[t], NULL, perm_task_run, &tasks[t]);
        }
        for (int t = 0; t < num_threads; ++t) pthread_join(threads[t], NULL);
        free(tasks); free(threads);
    }
    /* z (prover.rs:303-320): num_chunks zeros, a one, then the running product over (row in hypercube order, chunk) */
    const size_t flat = num_chunks << num_vars;
    ofe_t *z = (ofe_t *)calloc(flat, sizeof(ofe_t));
    uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * n);
    oracle_bh_iter(num_vars, order);
    size_t pos = num_chunks;
    ofe_t state = one;
    if (pos < flat) z[pos++] = one;
    for (size_t nth = 1; nth < n && pos < flat; ++nth)
        for (size_t chunk = 0; chunk < num_chunks && pos < flat; ++chunk) {
            mont_mul(FR, state.l, products[chunk * n + order[nth]].l, state.l);
            z[pos++] = state;
        }
    /* into_bh_order (prover.rs:333-341): poly_offset[b] = z[offset + num_chunks * nth_map[b]] */
    for (size_t nth = 0; nth < n; ++nth)
        for (size_t offset = 0; offset < num_chunks; ++offset) out[offset * n + order[nth]] = z[offset + num_chunks * nth];
    free(products); free(z); free(order);
}
void oracle_permutation_z_polys(size_t num_chunks, const ofe_t *const *values, const ofe_t *const *sigmas, size_t count, size_t num_vars,
                                const ofe_t *beta, const ofe_t *gamma, ofe_t *out) {
    oracle_permutation_z_polys_mt(num_chunks, values, sigmas, count, num_vars, beta, gamma, 1, out);
}
