// TEST: the C++ mirror of the reference interface (include/plonkish_cuda.hpp) against the oracle, bit-exact.
// Reads like the reference's own PCS test (pcs/multilinear.rs:293-333: setup -> commit -> open) plus the msm.rs
// edge behaviour.  Links libplonkish_cuda.so (product) and liboracle_bn254.so (checker).
#include "../../include/plonkish_cuda.hpp"
#include "../../oracle/bn254_oracle.h"

#include <cstdio>
#include <cstdlib>

using namespace plonkish;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t next_u64() {  // splitmix64
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static Fr random_fr() {  // a valid Montgomery representation: top limb below r's top limb
    Fr f;
    for (int i = 0; i < 3; ++i) f.l[i] = next_u64();
    f.l[3] = next_u64() % 0x30644E72E131A029ull;
    return f;
}
static std::vector<Fr> random_frs(size_t n) {
    std::vector<Fr> v(n);
    for (auto &f : v) f = random_fr();
    return v;
}
#define REQUIRE(cond)                                                      \
    do {                                                                   \
        if (!(cond)) {                                                     \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                      \
        }                                                                  \
    } while (0)

static G1Affine oracle_msm(const std::vector<Fr> &s, const std::vector<G1Affine> &b) {
    og1_jac_t jac;
    og1_affine_t out;
    oracle_variable_base_msm((const ofe_t *)s.data(), (const og1_affine_t *)b.data(), s.size(), 4, &jac);
    oracle_g1_to_affine(&jac, &out);
    G1Affine r;
    memcpy(&r, &out, sizeof(r));
    return r;
}
static Fr fr_sub(const Fr &a, const Fr &b) { Fr o; oracle_fe_sub(1, (const ofe_t *)&a, (const ofe_t *)&b, (ofe_t *)&o); return o; }
static Fr fr_add(const Fr &a, const Fr &b) { Fr o; oracle_fe_add(1, (const ofe_t *)&a, (const ofe_t *)&b, (ofe_t *)&o); return o; }
static Fr fr_mul(const Fr &a, const Fr &b) { Fr o; oracle_fe_mul(1, (const ofe_t *)&a, (const ofe_t *)&b, (ofe_t *)&o); return o; }
static Fr fr_inv(const Fr &a) { Fr o; oracle_fe_inv(1, (const ofe_t *)&a, (ofe_t *)&o); return o; }
static Fr fr_from(uint64_t v) { uint64_t c[4] = {v, 0, 0, 0}; Fr o; oracle_fe_from_canonical(1, c, (ofe_t *)&o); return o; }
// Lagrange interpolation through (i, msg[i]) at x: what Evaluations::evaluate computes (eval.rs:50-52)
static Fr interpolate(const std::vector<Fr> &msg, const Fr &x) {
    Fr total = fr_from(0);
    for (size_t j = 0; j < msg.size(); ++j) {
        Fr num = fr_from(1), den = fr_from(1);
        for (size_t i = 0; i < msg.size(); ++i) {
            if (i == j) continue;
            num = fr_mul(num, fr_sub(x, fr_from(i)));
            den = fr_mul(den, fr_sub(fr_from(j), fr_from(i)));
        }
        total = fr_add(total, fr_mul(msg[j], fr_mul(num, fr_inv(den))));
    }
    return total;
}

static int run();
int main() {
    try {
        return run();
    } catch (const std::exception &e) {  // no device / library error: the mirror has no CPU path
        fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
}
static int run() {
    init();
    og1_affine_t g_o;
    oracle_g1_generator(&g_o);
    G1Affine g;
    memcpy(&g, &g_o, sizeof(g));

    // ---- variable_base_msm (msm.rs:84-115): slices, a resident slice, the length assert, n = 0
    const size_t n = 3000;
    std::vector<G1Affine> bases(n);
    {
        uint64_t a[4] = {3, 0, 0, 0}, d[4] = {5, 0, 0, 0};
        oracle_known_dlog_bases(a, d, n, 4, (og1_affine_t *)bases.data());
    }
    auto scalars = random_frs(n);
    const G1Affine want = oracle_msm(scalars, bases);
    REQUIRE(variable_base_msm(scalars, bases) == want);
    {
        G1Bases resident(bases.data(), n);
        REQUIRE(variable_base_msm(scalars, resident) == want);
        REQUIRE(resident.to_host()[17] == bases[17]);
    }
    bool threw = false;
    try { variable_base_msm(std::vector<Fr>(scalars.begin(), scalars.begin() + 5), bases); } catch (const std::invalid_argument &) { threw = true; }
    REQUIRE(threw);  // assert_eq! at msm.rs:90
    REQUIRE(variable_base_msm(std::vector<Fr>{}, std::vector<G1Affine>{}).is_identity());  // documented deviation from msm.rs:154
    REQUIRE(fixed_base_msm(g, {fr_from(0), fr_from(1)})[1] == g);

    // ---- MultilinearKzg: setup -> commit / batch_commit -> open (pcs/multilinear.rs:293-333 shape)
    const size_t k = 7;
    auto ss = random_frs(k);
    auto pp = MultilinearKzgProverParam::setup(g, ss);
    REQUIRE(pp.num_vars() == k);
    std::vector<Fr> eq_scalars((size_t(2) << k) - 1);
    oracle_kzg_eq_scalars((const ofe_t *)ss.data(), k, (ofe_t *)eq_scalars.data());
    std::vector<std::vector<G1Affine>> eqs;
    for (size_t i = 0; i <= k; ++i) {
        std::vector<G1Affine> pts(size_t(1) << i);
        oracle_fixed_base_msm(&g_o, 5, (const ofe_t *)eq_scalars.data() + ((size_t(1) << i) - 1), pts.size(), 2, (og1_affine_t *)pts.data());
        REQUIRE(pp.eq(i).to_host() == pts);
        eqs.push_back(pts);
    }
    auto evals = random_frs(size_t(1) << k), evals2 = random_frs(size_t(1) << k);
    REQUIRE(pp.commit(evals) == oracle_msm(evals, eqs[k]));
    auto [comms, resident] = pp.batch_commit({&evals, &evals2});
    REQUIRE(comms[0] == oracle_msm(evals, eqs[k]) && comms[1] == oracle_msm(evals2, eqs[k]));
    REQUIRE(resident[1].evals() == evals2 && pp.commit(resident[0]) == comms[0]);
    auto point = random_frs(k);
    auto [q_comms, value] = pp.open(resident[0], point);
    std::vector<Fr> quotients(size_t(1) << k);
    Fr want_value;
    oracle_quotients((const ofe_t *)evals.data(), (const ofe_t *)point.data(), k, (ofe_t *)quotients.data(), (ofe_t *)&want_value);
    REQUIRE(value == want_value);
    for (size_t i = 0; i < k; ++i) {
        std::vector<Fr> q(quotients.begin() + (size_t(1) << i), quotients.begin() + (size_t(2) << i));
        REQUIRE(q_comms[i] == oracle_msm(q, eqs[i]));
    }
    threw = false;
    try { pp.commit(random_frs(size_t(2) << k)); } catch (const std::invalid_argument &) { threw = true; }
    REQUIRE(threw);  // "Too many variates of poly to commit" (pcs/multilinear.rs:26-58)
    {   // a ProverParam from host slices behaves the same
        MultilinearKzgProverParam from_host(eqs);
        REQUIRE(from_host.commit(evals) == comms[0]);
    }
    // g_prime merge (pcs/multilinear.rs:203-213)
    auto coeffs = random_frs(2);
    auto merged = linear_combination({&resident[0], &resident[1]}, coeffs);
    {
        const ofe_t *ps[2] = {(const ofe_t *)evals.data(), (const ofe_t *)evals2.data()};
        std::vector<Fr> w(evals.size());
        oracle_fr_linear_combination(ps, (const ofe_t *)coeffs.data(), 2, evals.size(), (ofe_t *)w.data());
        REQUIRE(merged.evals() == w);
    }

    // ---- Zeromorph / Gemini over the univariate SRS (zeromorph.rs:149-180, gemini.rs:98-128): kept quotients and folds,
    //      their commitments from sub-ranges, q_hat and f, the division by X - x
    {
        const size_t kk = 6, nn = size_t(1) << kk;
        const Fr s = random_fr();
        std::vector<Fr> pw(nn, fr_from(1));
        for (size_t i = 1; i < nn; ++i) pw[i] = fr_mul(pw[i - 1], s);
        std::vector<G1Affine> powers(nn);
        oracle_fixed_base_msm(&g_o, 5, (const ofe_t *)pw.data(), nn, 2, (og1_affine_t *)powers.data());
        G1Bases srs = univariate_setup(g, s, nn);
        REQUIRE(srs.to_host() == powers);
        auto f_evals = random_frs(nn), u = random_frs(kk);
        MultilinearPolynomial poly(f_evals.data(), nn);
        auto [q, value] = plonkish::quotients(poly, u);
        std::vector<Fr> want_q(nn);
        Fr want_value;
        oracle_quotients((const ofe_t *)f_evals.data(), (const ofe_t *)u.data(), kk, (ofe_t *)want_q.data(), (ofe_t *)&want_value);
        auto got_q = q.evals();
        REQUIRE(value == want_value && got_q[0] == fr_from(0));
        for (size_t i = 1; i < nn; ++i) REQUIRE(got_q[i] == want_q[i]);
        std::vector<size_t> sizes;
        for (size_t i = 0; i < kk; ++i) sizes.push_back(size_t(1) << i);
        auto q_comms = commit_packed(q, sizes, srs);
        for (size_t i = 0; i < kk; ++i) {
            std::vector<Fr> qi(got_q.begin() + sizes[i], got_q.begin() + 2 * sizes[i]);
            REQUIRE(q_comms[i] == oracle_msm(qi, std::vector<G1Affine>(powers.begin(), powers.begin() + sizes[i])));
        }
        auto ys = random_frs(kk), ss_ = random_frs(kk);
        const Fr z = random_fr(), c0 = random_fr();
        auto q_hat = zeromorph_q_hat(q, ys);
        auto ff = zeromorph_f(poly, q_hat, q, z, c0, ss_);
        std::vector<Fr> want_hat(nn, fr_from(0)), want_f(nn);
        for (size_t m = 0; m < nn; ++m) want_f[m] = fr_mul(z, f_evals[m]);
        for (size_t i = 0; i < kk; ++i) {
            for (size_t j = 0; j < sizes[i]; ++j) {
                want_hat[nn - sizes[i] + j] = fr_add(want_hat[nn - sizes[i] + j], fr_mul(ys[i], got_q[sizes[i] + j]));
                want_f[j] = fr_add(want_f[j], fr_mul(ss_[i], got_q[sizes[i] + j]));
            }
        }
        for (size_t m = 0; m < nn; ++m) want_f[m] = fr_add(want_f[m], want_hat[m]);
        want_f[0] = fr_add(want_f[0], c0);
        REQUIRE(q_hat.evals() == want_hat && ff.evals() == want_f);
        // div_rem by (X - x): quotient * (X - x) + remainder gives the polynomial back, the remainder is its value at x
        const Fr x = random_fr();
        auto [quot, rem] = div_linear(ff, x);
        auto qc = quot.evals();
        REQUIRE(qc[nn - 1] == fr_from(0));
        for (size_t m = 0; m < nn; ++m) REQUIRE(fr_sub(m ? qc[m - 1] : rem, fr_mul(x, qc[m])) == want_f[m]);
        // Gemini's folds: f_i[j] = (f_(i-1)[2j+1] - f_(i-1)[2j]) * u_(i-1) + f_(i-1)[2j], packed at offset 2^(kk-i); one as a slice
        auto folds = gemini_folds(poly, u);
        auto packed = folds.evals();
        std::vector<Fr> cur = f_evals;
        for (size_t i = 1; i < kk; ++i) {
            std::vector<Fr> nxt(cur.size() / 2);
            for (size_t j = 0; j < nxt.size(); ++j) nxt[j] = fr_add(fr_mul(fr_sub(cur[2 * j + 1], cur[2 * j]), u[i - 1]), cur[2 * j]);
            for (size_t j = 0; j < nxt.size(); ++j) REQUIRE(packed[nxt.size() + j] == nxt[j]);
            cur = nxt;
        }
        auto f2 = slice(folds, nn >> 2, nn >> 2);
        REQUIRE(variable_base_msm(f2, srs) == commit_packed(folds, {nn >> 2}, srs)[0]);
    }

    // ---- ClassicSumCheck::prove on a zero check eq * (a*b - c), c = a o b: sum 0 (classic.rs:208-240)
    {
        const size_t kk = 6, nn = size_t(1) << kk;
        auto a = random_frs(nn), b = random_frs(nn), y = random_frs(kk);
        std::vector<Fr> c(nn), eq(1, fr_from(1));
        for (size_t i = 0; i < nn; ++i) c[i] = fr_mul(a[i], b[i]);
        for (size_t i = 0; i < kk; ++i) {  // eq(x, y) table, lowest variable first
            std::vector<Fr> next(eq.size() * 2);
            for (size_t j = 0; j < eq.size(); ++j) {
                next[eq.size() + j] = fr_mul(eq[j], y[i]);
                next[j] = fr_sub(eq[j], next[eq.size() + j]);
            }
            eq.swap(next);
        }
        MultilinearPolynomial p_eq(eq.data(), nn), p_a(a.data(), nn), p_b(b.data(), nn), p_c(c.data(), nn);
        const Fr one = fr_from(1), minus_one = fr_sub(fr_from(0), one);
        std::vector<std::vector<Fr>> msgs;
        auto [challenges, finals] = sum_check_prove(
            {&p_eq, &p_a, &p_b, &p_c}, {{one, {1, 2}}, {minus_one, {3}}}, 0, fr_from(0),
            [&](const std::vector<Fr> &msg) { msgs.push_back(msg); return random_fr(); }, interpolate, fr_sub);
        REQUIRE(challenges.size() == kk && msgs.size() == kk && msgs[0].size() == 4);
        Fr claim = fr_from(0);
        for (size_t r = 0; r < kk; ++r) {  // the verifier's consistency check (classic.rs:168-190)
            REQUIRE(fr_add(msgs[r][0], msgs[r][1]) == claim);
            claim = interpolate(msgs[r], challenges[r]);
        }
        REQUIRE(claim == fr_mul(finals[0], fr_sub(fr_mul(finals[1], finals[2]), finals[3])));
        // the first message against the oracle's round
        const ofe_t *ps[4] = {(const ofe_t *)eq.data(), (const ofe_t *)a.data(), (const ofe_t *)b.data(), (const ofe_t *)c.data()};
        const Fr cf[2] = {one, minus_one};
        const uint32_t offs[3] = {0, 2, 3}, fl[3] = {1, 2, 3};
        Fr r0[3];
        oracle_sumcheck_round(ps, 4, nn / 2, (const ofe_t *)cf, offs, fl, 2, 0, 3, (ofe_t *)r0);
        REQUIRE(r0[0] == msgs[0][1] && r0[1] == msgs[0][2] && r0[2] == msgs[0][3]);
    }
    plonkish_cuda_shutdown();
    printf("cpp mirror ok\n");
    return 0;
}
