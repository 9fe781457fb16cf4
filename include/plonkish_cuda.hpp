// plonkish_cuda.hpp — C++ host-side mirror of the reference interface over the C ABI (plonkish_cuda.h).
//
// The reference is Rust; its toolchain is not in this image, so the compiled host side above the C ABI is this
// header (the Rust binding a maintainer would add is bindings/plonkish_cuda/src/lib.rs; the Python mirror used by
// the test-suite is plonkish_b200/).  Names, argument meaning and error behaviour follow the reference
// (paths relative to /root/reference/plonkish_backend/src):
//
//   variable_base_msm                     util/arithmetic/msm.rs:84-115   (+ the callers' to_affine())
//   fixed_base_msm                        util/arithmetic/msm.rs:16-31, 67-81 (+ batch_normalize)
//   MultilinearKzgProverParam::{num_vars, eq}   pcs/multilinear/kzg.rs:55-77
//   MultilinearKzg::{setup (prover half), commit, batch_commit, open}   pcs/multilinear/kzg.rs:167-212, 252-302
//   MultilinearPolynomial (resident evaluations), linear_combination   pcs/multilinear.rs:203-213
//   ClassicSumCheck::prove                piop/sum_check/classic.rs:208-240
//
// The reference panics on its own errors (assert_eq! at msm.rs:90, index panic at msm.rs:154) and has no
// fallback; here every non-zero return code of the C ABI throws plonkish::Error.  There is no CPU path.
#pragma once
#include "plonkish_cuda.h"

#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace plonkish {

struct Fr {        // bn256::Fr: 4 x u64 little-endian limbs, Montgomery form
    uint64_t l[4];
    bool operator==(const Fr &o) const { return std::memcmp(l, o.l, sizeof(l)) == 0; }
};
struct G1Affine {  // bn256::G1Affine: x || y Montgomery Fq, (0, 0) = identity
    uint64_t x[4], y[4];
    bool operator==(const G1Affine &o) const { return std::memcmp(this, &o, sizeof(*this)) == 0; }
    bool is_identity() const {
        for (int i = 0; i < 4; ++i)
            if (x[i] | y[i]) return false;
        return true;
    }
};
static_assert(sizeof(Fr) == PLONKISH_CUDA_SCALAR_BYTES && sizeof(G1Affine) == PLONKISH_CUDA_AFFINE_BYTES, "layout of the C ABI");

struct Error : std::runtime_error {
    int code;
    Error(const std::string &what, int rc) : std::runtime_error(what + " failed (" + std::to_string(rc) + "): " + plonkish_cuda_last_error()), code(rc) {}
};
inline void check(int rc, const char *what) {
    if (rc != PLONKISH_CUDA_OK) throw Error(what, rc);
}
inline void init(int n_devices = 0) { check(plonkish_cuda_init(n_devices), "plonkish_cuda_init"); }

// A base slice resident on the GPU: what a ProverParam owns (eqs[k], kzg.rs:74-76; powers_of_s_g1, univariate/kzg.rs:24-30).
class G1Bases {
  public:
    G1Bases() = default;
    G1Bases(const G1Affine *bases, size_t n, int device = 0) : n_(n) { check(plonkish_cuda_bases_register(device, bases, n, &handle_), "plonkish_cuda_bases_register"); }
    static G1Bases adopt(uint64_t handle, size_t n) { G1Bases b; b.handle_ = handle; b.n_ = n; return b; }
    G1Bases(G1Bases &&o) noexcept : handle_(o.handle_), n_(o.n_) { o.handle_ = 0; }
    G1Bases &operator=(G1Bases &&o) noexcept { release(); handle_ = o.handle_; n_ = o.n_; o.handle_ = 0; return *this; }
    G1Bases(const G1Bases &) = delete;
    G1Bases &operator=(const G1Bases &) = delete;
    ~G1Bases() { release(); }
    uint64_t handle() const { return handle_; }
    size_t len() const { return n_; }
    std::vector<G1Affine> to_host() const {
        std::vector<G1Affine> out(n_);
        check(plonkish_cuda_bases_read(handle_, 0, n_, out.data()), "plonkish_cuda_bases_read");
        return out;
    }

  private:
    void release() { if (handle_) plonkish_cuda_bases_release(handle_); handle_ = 0; }
    uint64_t handle_ = 0;
    size_t n_ = 0;
};

// poly.evals() kept in HBM between commit and open.
class MultilinearPolynomial {
  public:
    MultilinearPolynomial() = default;
    MultilinearPolynomial(const Fr *evals, size_t n, int device = 0) : n_(n) { check(plonkish_cuda_scalars_register(device, evals, n, &handle_), "plonkish_cuda_scalars_register"); }
    static MultilinearPolynomial adopt(uint64_t handle, size_t n) { MultilinearPolynomial p; p.handle_ = handle; p.n_ = n; return p; }
    MultilinearPolynomial(MultilinearPolynomial &&o) noexcept : handle_(o.handle_), n_(o.n_) { o.handle_ = 0; }
    MultilinearPolynomial &operator=(MultilinearPolynomial &&o) noexcept { release(); handle_ = o.handle_; n_ = o.n_; o.handle_ = 0; return *this; }
    MultilinearPolynomial(const MultilinearPolynomial &) = delete;
    MultilinearPolynomial &operator=(const MultilinearPolynomial &) = delete;
    ~MultilinearPolynomial() { release(); }
    uint64_t handle() const { return handle_; }
    size_t len() const { return n_; }
    size_t num_vars() const { size_t k = 0; while ((size_t(1) << k) < n_) ++k; return k; }
    std::vector<Fr> evals() const {
        std::vector<Fr> out(n_);
        check(plonkish_cuda_scalars_read(handle_, 0, n_, out.data()), "plonkish_cuda_scalars_read");
        return out;
    }

  private:
    void release() { if (handle_) plonkish_cuda_scalars_release(handle_); handle_ = 0; }
    uint64_t handle_ = 0;
    size_t n_ = 0;
};

// msm.rs:84-115 followed by the callers' to_affine().  A length mismatch is the reference's assert_eq! (msm.rs:90).
inline G1Affine variable_base_msm(const std::vector<Fr> &scalars, const std::vector<G1Affine> &bases) {
    if (scalars.size() != bases.size()) throw std::invalid_argument("variable_base_msm: scalars and bases differ in length (msm.rs:90)");
    G1Affine out{};
    check(plonkish_cuda_msm_bn254_g1(scalars.data(), bases.data(), 0, scalars.size(), &out), "plonkish_cuda_msm_bn254_g1");
    return out;
}
inline G1Affine variable_base_msm(const std::vector<Fr> &scalars, const G1Bases &bases) {
    if (scalars.size() > bases.len()) throw std::invalid_argument("variable_base_msm: more scalars than registered bases (msm.rs:90)");
    G1Affine out{};
    check(plonkish_cuda_msm_bn254_g1(scalars.data(), nullptr, bases.handle(), scalars.size(), &out), "plonkish_cuda_msm_bn254_g1");
    return out;
}
inline G1Affine variable_base_msm(const MultilinearPolynomial &poly, const G1Bases &bases) {
    if (poly.len() > bases.len()) throw std::invalid_argument("variable_base_msm: more scalars than registered bases (msm.rs:90)");
    G1Affine out{};
    check(plonkish_cuda_msm_bn254_g1_resident(poly.handle(), bases.handle(), poly.len(), &out), "plonkish_cuda_msm_bn254_g1_resident");
    return out;
}
// msm.rs:16-31, 67-81 + batch_normalize: out[i] = scalars[i] * base.
inline std::vector<G1Affine> fixed_base_msm(const G1Affine &base, const std::vector<Fr> &scalars, int device = 0) {
    std::vector<G1Affine> out(scalars.size());
    check(plonkish_cuda_fixed_base_msm_bn254_g1(device, &base, scalars.data(), scalars.size(), out.data()), "plonkish_cuda_fixed_base_msm_bn254_g1");
    return out;
}
// The G1 half of UnivariateKzg::setup (pcs/univariate/kzg.rs:175-195): powers_of_s_g1[i] = s^i * g1, i < poly_size,
// built and kept on the GPU; commit_coeffs (univariate/kzg.rs:24-30) is variable_base_msm(coeffs, powers_of_s_g1).
inline G1Bases univariate_setup(const G1Affine &g1, const Fr &s, size_t poly_size, int device = 0) {
    uint64_t h = 0;
    check(plonkish_cuda_kzg_setup_powers_bn254(device, &g1, &s, poly_size, &h), "plonkish_cuda_kzg_setup_powers_bn254");
    return G1Bases::adopt(h, poly_size);
}
// MultilinearPolynomial::eq_xy(y): eq(x, y) over the hypercube, resident (the zero-check factor, classic.rs:57-61).
inline MultilinearPolynomial eq_xy(const std::vector<Fr> &y, int device = 0) {
    uint64_t h = 0;
    check(plonkish_cuda_eq_table(device, y.data(), y.size(), &h), "plonkish_cuda_eq_table");
    return MultilinearPolynomial::adopt(h, size_t(1) << y.size());
}
// pcs/multilinear.rs:203-213: sum_i coeffs[i] * polys[i].
inline MultilinearPolynomial linear_combination(const std::vector<const MultilinearPolynomial *> &polys, const std::vector<Fr> &coeffs) {
    if (polys.empty() || polys.size() != coeffs.size()) throw std::invalid_argument("linear_combination: polys and coeffs differ in length");
    std::vector<uint64_t> hs;
    for (auto *p : polys) hs.push_back(p->handle());
    uint64_t h = 0;
    check(plonkish_cuda_fr_linear_combination(hs.data(), coeffs.data(), hs.size(), polys[0]->len(), &h), "plonkish_cuda_fr_linear_combination");
    return MultilinearPolynomial::adopt(h, polys[0]->len());
}

// ---- the polynomial side of Zeromorph / Gemini over the univariate SRS (pcs/multilinear/zeromorph.rs, gemini.rs) ----
// `quotients(poly, point, ..)` (pcs/multilinear.rs:72-107) kept in HBM: quotient i (2^i values) at element offset 2^i of the
// returned vector (element 0 zero), and the remainder poly(point).
inline std::pair<MultilinearPolynomial, Fr> quotients(const MultilinearPolynomial &poly, const std::vector<Fr> &point) {
    if ((size_t(1) << point.size()) != poly.len()) throw std::invalid_argument("quotients: the point does not match the polynomial (multilinear.rs:77)");
    uint64_t h = 0;
    Fr eval{};
    check(plonkish_cuda_fr_quotients(poly.handle(), point.data(), point.size(), &h, &eval), "plonkish_cuda_fr_quotients");
    return {MultilinearPolynomial::adopt(h, poly.len()), eval};
}
// The folds fs[1..] of Gemini::open (gemini.rs:98-108), packed: f_i (2^(k-i) values) at element offset 2^(k-i).
inline MultilinearPolynomial gemini_folds(const MultilinearPolynomial &poly, const std::vector<Fr> &point) {
    if (point.empty() || (size_t(1) << point.size()) != poly.len()) throw std::invalid_argument("gemini_folds: the point does not match the polynomial");
    uint64_t h = 0;
    check(plonkish_cuda_fr_gemini_folds(poly.handle(), point.data(), point.size(), &h), "plonkish_cuda_fr_gemini_folds");
    return MultilinearPolynomial::adopt(h, poly.len());
}
// UnivariateKzg::batch_commit over packed quotients / folds (zeromorph.rs:150, gemini.rs:124-128): the parts of `packed` at
// element offset sizes[j], sizes[j] values long, each against powers_of_s_g1[..sizes[j]] (commit_coeffs, univariate/kzg.rs:24-30).
inline std::vector<G1Affine> commit_packed(const MultilinearPolynomial &packed, const std::vector<size_t> &sizes, const G1Bases &powers_of_s_g1) {
    std::vector<uint64_t> hs(sizes.size(), powers_of_s_g1.handle());
    std::vector<G1Affine> out(sizes.size());
    if (!sizes.empty())
        check(plonkish_cuda_msm_bn254_g1_many_resident(packed.handle(), sizes.data(), hs.data(), sizes.data(), sizes.size(), out.data()),
              "plonkish_cuda_msm_bn254_g1_many_resident");
    return out;
}
// A polynomial on a sub-range of a resident vector (shares the memory): one fold or one quotient on its own.
inline MultilinearPolynomial slice(const MultilinearPolynomial &v, size_t offset, size_t n) {
    uint64_t h = 0;
    check(plonkish_cuda_scalars_slice(v.handle(), offset, n, &h), "plonkish_cuda_scalars_slice");
    return MultilinearPolynomial::adopt(h, n);
}
// q_hat (zeromorph.rs:157-168) and f (zeromorph.rs:175-180) of Zeromorph::open from the packed quotients.
inline MultilinearPolynomial zeromorph_q_hat(const MultilinearPolynomial &q, const std::vector<Fr> &powers_of_y) {
    uint64_t h = 0;
    check(plonkish_cuda_zeromorph_q_hat_bn254(q.handle(), powers_of_y.data(), powers_of_y.size(), &h), "plonkish_cuda_zeromorph_q_hat_bn254");
    return MultilinearPolynomial::adopt(h, q.len());
}
inline MultilinearPolynomial zeromorph_f(const MultilinearPolynomial &poly, const MultilinearPolynomial &q_hat, const MultilinearPolynomial &q, const Fr &z,
                                         const Fr &c0, const std::vector<Fr> &q_scalars) {
    uint64_t h = 0;
    check(plonkish_cuda_zeromorph_f_bn254(poly.handle(), q_hat.handle(), q.handle(), &z, &c0, q_scalars.data(), q_scalars.size(), &h),
          "plonkish_cuda_zeromorph_f_bn254");
    return MultilinearPolynomial::adopt(h, poly.len());
}
// UnivariatePolynomial::div_rem by (X - z) (poly/univariate.rs:144-168): (quotient, remainder = value at z).
inline std::pair<MultilinearPolynomial, Fr> div_linear(const MultilinearPolynomial &coeffs, const Fr &z) {
    uint64_t h = 0;
    Fr rem{};
    check(plonkish_cuda_fr_div_linear(coeffs.handle(), &z, &h, &rem), "plonkish_cuda_fr_div_linear");
    return {MultilinearPolynomial::adopt(h, coeffs.len()), rem};
}

// MultilinearKzgProverParams { g1, eqs } (kzg.rs:55-77) with every eqs[k] resident.
class MultilinearKzgProverParam {
  public:
    // The prover half of MultilinearKzg::setup (kzg.rs:167-212) for the trapdoor ss, built on the device.
    static MultilinearKzgProverParam setup(const G1Affine &g1, const std::vector<Fr> &ss, int device = 0) {
        std::vector<uint64_t> hs(ss.size() + 1);
        check(plonkish_cuda_kzg_setup_eqs_bn254(device, &g1, ss.data(), ss.size(), hs.data()), "plonkish_cuda_kzg_setup_eqs_bn254");
        MultilinearKzgProverParam pp;
        for (size_t k = 0; k < hs.size(); ++k) pp.eqs_.push_back(G1Bases::adopt(hs[k], size_t(1) << k));
        return pp;
    }
    // From host slices (a parameter file): eqs[k] holds 2^k bases.
    explicit MultilinearKzgProverParam(const std::vector<std::vector<G1Affine>> &eqs, int device = 0) {
        for (size_t k = 0; k < eqs.size(); ++k) {
            if (eqs[k].size() != (size_t(1) << k)) throw std::invalid_argument("MultilinearKzgProverParam: eqs[k] must hold 2^k bases");
            eqs_.emplace_back(eqs[k].data(), eqs[k].size(), device);
        }
    }
    size_t num_vars() const { return eqs_.size() - 1; }           // kzg.rs:68-70
    const G1Bases &eq(size_t num_vars) const { return eqs_.at(num_vars); }  // kzg.rs:74-76

    // kzg.rs:252-257
    G1Affine commit(const std::vector<Fr> &evals) const { return variable_base_msm(evals, eq(checked_num_vars(evals.size(), "commit"))); }
    G1Affine commit(const MultilinearPolynomial &poly) const { return variable_base_msm(poly, eq(checked_num_vars(poly.len(), "commit"))); }
    // kzg.rs:259-274; polynomials of one size, left resident for the later open
    std::pair<std::vector<G1Affine>, std::vector<MultilinearPolynomial>> batch_commit(const std::vector<const std::vector<Fr> *> &polys) const {
        if (polys.empty()) return {};
        const size_t n = polys[0]->size();
        const size_t k = checked_num_vars(n, "batch commit");
        std::vector<const void *> ptrs;
        for (auto *p : polys) {
            if (p->size() != n) throw std::invalid_argument("batch_commit: polynomials of different sizes go down one by one");
            ptrs.push_back(p->data());
        }
        std::vector<G1Affine> comms(polys.size());
        std::vector<uint64_t> hs(polys.size());
        check(plonkish_cuda_msm_bn254_g1_batch_keep(ptrs.data(), ptrs.size(), eq(k).handle(), n, comms.data(), hs.data()), "plonkish_cuda_msm_bn254_g1_batch_keep");
        std::vector<MultilinearPolynomial> resident;
        for (uint64_t h : hs) resident.push_back(MultilinearPolynomial::adopt(h, n));
        return {std::move(comms), std::move(resident)};
    }
    // kzg.rs:276-302: the quotient commitments in transcript order (kzg.rs:299) and f(point) (the remainder, kzg.rs:295)
    std::pair<std::vector<G1Affine>, Fr> open(const MultilinearPolynomial &poly, const std::vector<Fr> &point) const {
        if ((size_t(1) << point.size()) != poly.len()) throw std::invalid_argument("open: point and polynomial differ in the number of variables (multilinear.rs:77)");
        checked_num_vars(poly.len(), "open");
        std::vector<uint64_t> hs;
        for (size_t i = 0; i < point.size(); ++i) hs.push_back(eq(i).handle());
        std::vector<G1Affine> comms(point.size());
        Fr eval{};
        check(plonkish_cuda_kzg_open_bn254(poly.handle(), hs.data(), point.data(), point.size(), comms.data(), &eval), "plonkish_cuda_kzg_open_bn254");
        return {std::move(comms), eval};
    }

  private:
    MultilinearKzgProverParam() = default;
    size_t checked_num_vars(size_t n, const char *function) const {  // validate_input, pcs/multilinear.rs:26-58
        size_t k = 0;
        while ((size_t(1) << k) < n) ++k;
        if ((size_t(1) << k) != n) throw std::invalid_argument("a multilinear polynomial has 2^k evaluations");
        if (k > num_vars())
            throw std::invalid_argument(std::string("Too many variates of poly to ") + function + " (param supports variates up to " + std::to_string(num_vars()) +
                                        " but got " + std::to_string(k) + ")");
        return k;
    }
    std::vector<G1Bases> eqs_;
};

// One flattened term of a sum-check expression: coeff * prod polys[factors].
struct SumCheckTerm {
    Fr coeff;
    std::vector<uint32_t> factors;
};
// ClassicSumCheck::prove (piop/sum_check/classic.rs:208-240).  `squeeze(msg)` stands for msg.write(transcript) +
// transcript.squeeze_challenge() (classic.rs:226-229), `evaluate(msg, challenge)` for msg.evaluate (eval.rs:50-52);
// `sub(a, b)` is field subtraction for evals[0] = sum - evals[1] (eval.rs:128).  Returns (challenges, evals).
inline std::pair<std::vector<Fr>, std::vector<Fr>> sum_check_prove(const std::vector<const MultilinearPolynomial *> &polys, const std::vector<SumCheckTerm> &terms,
                                                                   int common_poly, Fr sum, const std::function<Fr(const std::vector<Fr> &)> &squeeze,
                                                                   const std::function<Fr(const std::vector<Fr> &, const Fr &)> &evaluate,
                                                                   const std::function<Fr(const Fr &, const Fr &)> &sub) {
    std::vector<uint64_t> hs;
    for (auto *p : polys) hs.push_back(p->handle());
    std::vector<Fr> coeffs;
    std::vector<uint32_t> offsets{0}, flat;
    for (auto &t : terms) {
        coeffs.push_back(t.coeff);
        flat.insert(flat.end(), t.factors.begin(), t.factors.end());
        offsets.push_back((uint32_t)flat.size());
    }
    if (flat.empty()) flat.push_back(0);
    const size_t num_vars = polys.at(0)->num_vars();
    uint64_t state = 0;
    check(plonkish_cuda_sumcheck_new(hs.data(), hs.size(), num_vars, coeffs.data(), offsets.data(), flat.data(), terms.size(), common_poly, &state), "plonkish_cuda_sumcheck_new");
    struct Guard { uint64_t s; ~Guard() { plonkish_cuda_sumcheck_free(s); } } guard{state};
    const int degree = plonkish_cuda_sumcheck_degree(state);
    if (degree < 0) throw Error("plonkish_cuda_sumcheck_degree", degree);
    std::vector<Fr> challenges;
    for (size_t round = 0; round < num_vars; ++round) {
        std::vector<Fr> msg(degree + 1);
        check(plonkish_cuda_sumcheck_round(state, msg.data() + 1), "plonkish_cuda_sumcheck_round");
        msg[0] = sub(sum, msg[1]);
        const Fr challenge = squeeze(msg);
        sum = evaluate(msg, challenge);
        check(plonkish_cuda_sumcheck_fix_var(state, &challenge), "plonkish_cuda_sumcheck_fix_var");
        challenges.push_back(challenge);
    }
    std::vector<Fr> evals(polys.size());
    check(plonkish_cuda_sumcheck_final_evals(state, evals.data()), "plonkish_cuda_sumcheck_final_evals");
    return {std::move(challenges), std::move(evals)};
}

}  // namespace plonkish
