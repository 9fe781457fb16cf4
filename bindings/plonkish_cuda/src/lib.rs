//! Rust side of the drop-in: `variable_base_msm` for `bn256::G1Affine` on a B200.
//!
//! Binds include/plonkish_cuda.h.  `plonkish_backend::util::arithmetic::msm::variable_base_msm`
//! (util/arithmetic/msm.rs:84-115) keeps its signature and routes here when
//! `C == bn256::G1Affine` (see INTEGRATION.md for the three-line patch).  There is no CPU
//! fallback: a non-zero return code panics, as the reference panics on its own errors
//! (msm.rs:90, :154).
use halo2_curves::bn256::{Fr, G1Affine, G1};
use halo2_curves::CurveAffine;
use std::{ffi::CStr, mem::size_of, os::raw::{c_char, c_int, c_void}, sync::{Once, OnceLock}};

extern "C" {
    fn plonkish_cuda_init(n_devices: c_int) -> c_int;
    fn plonkish_cuda_device_count() -> c_int;
    fn plonkish_cuda_last_error() -> *const c_char;
    fn plonkish_cuda_bases_register(device: c_int, bases: *const c_void, n: usize, handle: *mut u64) -> c_int;
    fn plonkish_cuda_bases_register_sharded(n_gpus: c_int, bases: *const c_void, n: usize, handle: *mut u64) -> c_int;
    fn plonkish_cuda_bases_cached(device: c_int, bases: *const c_void, n: usize, handle: *mut u64) -> c_int;
    fn plonkish_cuda_bases_cached_sharded(n_gpus: c_int, bases: *const c_void, n: usize, handle: *mut u64) -> c_int;
    fn plonkish_cuda_bases_cache_evict(bases: *const c_void) -> c_int;
    fn plonkish_cuda_bases_cache_limit(max_bytes: usize) -> c_int;
    fn plonkish_cuda_msm_bn254_g1(scalars: *const c_void, bases: *const c_void, handle: u64, n: usize, out: *mut c_void) -> c_int;
    fn plonkish_cuda_msm_bn254_g1_multi(n_gpus: c_int, scalars: *const c_void, bases: *const c_void, handle: u64, n: usize, out: *mut c_void) -> c_int;
    fn plonkish_cuda_msm_bn254_g1_gather(scalars: *const *const c_void, bases: *const *const c_void, n: usize, out: *mut c_void) -> c_int;
    fn plonkish_cuda_msm_bn254_g1_batch(scalars_list: *const *const c_void, count: usize, handle: u64, n: usize, out: *mut c_void) -> c_int;
    fn plonkish_cuda_msm_bn254_g1_many(scalars_list: *const *const c_void, handles: *const u64, ns: *const usize, count: usize, out: *mut c_void) -> c_int;
    fn plonkish_cuda_timer_config(mode: c_int, depth: c_int) -> c_int;
    fn plonkish_cuda_bases_release(handle: u64) -> c_int;
    fn plonkish_cuda_bases_read(handle: u64, offset: usize, n: usize, out: *mut c_void) -> c_int;
    fn plonkish_cuda_scalars_release(handle: u64) -> c_int;
    fn plonkish_cuda_msm_bn254_g1_batch_keep(scalars_list: *const *const c_void, count: usize, handle: u64, n: usize, out: *mut c_void, scalars_handles: *mut u64) -> c_int;
    fn plonkish_cuda_fr_linear_combination(handles: *const u64, coeffs: *const c_void, count: usize, n: usize, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_fr_div_linear(handle: u64, z: *const c_void, out_quotient: *mut u64, out_rem: *mut c_void) -> c_int;
    fn plonkish_cuda_fr_quotients(handle: u64, point: *const c_void, num_vars: usize, out_q: *mut u64, out_eval: *mut c_void) -> c_int;
    fn plonkish_cuda_scalars_slice(handle: u64, offset: usize, n: usize, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_fr_linear_combination_padded(handles: *const u64, coeffs: *const c_void, count: usize, n: usize, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_fr_gemini_folds(handle: u64, point: *const c_void, num_vars: usize, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_msm_bn254_g1_many_resident(scalars_handle: u64, offsets: *const usize, bases_handles: *const u64, ns: *const usize, count: usize,
                                                out: *mut c_void) -> c_int;
    fn plonkish_cuda_zeromorph_q_hat_bn254(q_handle: u64, weights: *const c_void, num_vars: usize, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_zeromorph_f_bn254(poly: u64, q_hat: u64, q: u64, z: *const c_void, c0: *const c_void, q_scalars: *const c_void, num_vars: usize,
                                       out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_fr_affine_table(device: c_int, num_vars: usize, polys: *const u64, rotations: *const i32, coeffs: *const c_void, count: usize,
                                     constant: *const c_void, identity_coeff: *const c_void, sparse_rows: *const u64, sparse_values: *const c_void,
                                     sparse_count: usize, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_fr_expression_table(polys: *const u64, num_polys: usize, num_vars: usize, coeffs: *const c_void, offsets: *const u32, term_polys: *const u32,
                                         num_terms: usize, common: c_int, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_lookup_m_poly_bn254(input: u64, table: u64, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_lookup_h_poly_bn254(input: u64, table: u64, m: u64, gamma: *const c_void, out_handle: *mut u64) -> c_int;
    fn plonkish_cuda_fr_evaluate(handle: u64, points: *const c_void, num_vars: usize, count: usize, out_evals: *mut c_void) -> c_int;
    fn plonkish_cuda_scalars_register(device: c_int, scalars: *const c_void, n: usize, handle: *mut u64) -> c_int;
    fn plonkish_cuda_msm_bn254_g1_resident(scalars_handle: u64, bases_handle: u64, n: usize, out: *mut c_void) -> c_int;
    fn plonkish_cuda_permutation_z_polys_bn254(values: *const u64, sigmas: *const u64, count: usize, num_chunks: usize, num_vars: usize,
                                               beta: *const c_void, gamma: *const c_void, out_handles: *mut u64) -> c_int;
    fn plonkish_cuda_kzg_open_bn254(scalars_handle: u64, eq_handles: *const u64, point: *const c_void, num_vars: usize, out_comms: *mut c_void, out_eval: *mut c_void) -> c_int;
    fn plonkish_cuda_fixed_base_msm_bn254_g1(device: c_int, base: *const c_void, scalars: *const c_void, n: usize, out: *mut c_void) -> c_int;
    fn plonkish_cuda_kzg_setup_eqs_bn254(device: c_int, g1: *const c_void, ss: *const c_void, num_vars: usize, handles_out: *mut u64) -> c_int;
    fn plonkish_cuda_kzg_setup_powers_bn254(device: c_int, g1: *const c_void, s: *const c_void, n: usize, handle: *mut u64) -> c_int;
    fn plonkish_cuda_eq_table(device: c_int, y: *const c_void, num_vars: usize, handle: *mut u64) -> c_int;
    fn plonkish_cuda_sumcheck_new(polys: *const u64, num_polys: usize, num_vars: usize, coeffs: *const c_void, offsets: *const u32, term_polys: *const u32,
                                  num_terms: usize, common_poly: c_int, state: *mut u64) -> c_int;
    fn plonkish_cuda_sumcheck_degree(state: u64) -> c_int;
    fn plonkish_cuda_sumcheck_round(state: u64, out_evals: *mut c_void) -> c_int;
    fn plonkish_cuda_sumcheck_round_factored(state: u64, out_evals: *mut c_void) -> c_int;
    fn plonkish_cuda_sumcheck_fix_var(state: u64, challenge: *const c_void) -> c_int;
    fn plonkish_cuda_sumcheck_final_evals(state: u64, out_evals: *mut c_void) -> c_int;
    fn plonkish_cuda_sumcheck_free(state: u64) -> c_int;
}

static INIT: Once = Once::new();
/// Slices shorter than this are not worth a resident copy: they are the throw-away vectors of
/// sum_with_scalar folds (pcs.rs:175), not ProverParam slices.  Resident slices get the table of
/// window multiples (no doubling chain: 0.5 ms instead of 1.7 ms for a small MSM).
const REGISTER_MIN: usize = 1 << 8;

fn check(rc: c_int, what: &str) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(plonkish_cuda_last_error()) }.to_string_lossy().into_owned();
        panic!("{what} failed ({rc}): {msg}");
    }
}

fn init() {
    INIT.call_once(|| {
        // halo2_curves types are not repr(C)-guaranteed: probe the layout once.
        assert_eq!(size_of::<Fr>(), 32);
        assert_eq!(size_of::<G1Affine>(), 64);
        let g = G1Affine::generator();
        let bytes: [u8; 64] = unsafe { std::mem::transmute(g) };
        // x = 1 and y = 2 in Montgomery form: R mod p and 2R mod p, little-endian limbs.
        assert_eq!(&bytes[..8], &0xd35d438dc58f0d9du64.to_le_bytes());
        assert_eq!(&bytes[32..40], &0xa6ba871b8b1e1b3au64.to_le_bytes());
        check(unsafe { plonkish_cuda_init(0) }, "plonkish_cuda_init");
        if let Some(bytes) = std::env::var("PLONKISH_CUDA_CACHE_BYTES").ok().and_then(|v| v.parse::<usize>().ok()) {
            check(unsafe { plonkish_cuda_bases_cache_limit(bytes) }, "plonkish_cuda_bases_cache_limit");
        }
        // The per-call timer line stays on the Rust side (msm.rs:92 is kept by the patch); the library prints the lines
        // of the composite entry points (batch / many / open), which replace several reference calls each.
        if cfg!(feature = "timer") {
            check(unsafe { plonkish_cuda_timer_config(2, timer_depth()) }, "plonkish_cuda_timer_config");
        }
    });
}

/// Nesting depth the library's own timer lines are printed at (PLONKISH_CUDA_TIMER_DEPTH; ark_std keeps its
/// indentation counter private): 2 inside HyperPlonk::prove's phase timers, the only place the composite
/// entry points are called from.
fn timer_depth() -> c_int {
    std::env::var("PLONKISH_CUDA_TIMER_DEPTH").ok().and_then(|v| v.parse().ok()).unwrap_or(2)
}

/// GPUs a single MSM is point-sharded over (PLONKISH_CUDA_GPUS, default 1): msm.rs:101-114 with GPUs for threads.
fn gpus() -> c_int {
    static GPUS: OnceLock<c_int> = OnceLock::new();
    *GPUS.get_or_init(|| {
        let want = std::env::var("PLONKISH_CUDA_GPUS").ok().and_then(|v| v.parse::<c_int>().ok()).unwrap_or(1);
        want.clamp(1, unsafe { plonkish_cuda_device_count() }.max(1))
    })
}
/// Below this many points a sharded MSM is slower than one GPU (launch chain + gather ~0.75 ms).
const MULTI_MIN: usize = 1 << 21;

/// Handle for a borrowed base slice.  The shim cannot know whether `bases` is a ProverParam's static SRS slice or a
/// temporary whose address the allocator will reuse (IPA's g_lo / g_hi, Hyrax rows, `trim`'s `to_vec()` of a second
/// setup), so it never trusts an address: `plonkish_cuda_bases_cached` keys by address but validates a hit against a
/// content fingerprint taken at registration, re-registers on mismatch, serves prefixes of a longer cached slice
/// (`&powers_of_s_g1[..len]`, univariate/kzg.rs:28) from the one registration, and evicts least-recently-used tables
/// beyond its byte limit.  Owners that know their lifetime use `RegisteredBases` instead.
fn cached_handle(bases: &[G1Affine], n_gpus: c_int) -> u64 {
    let mut h = 0u64;
    if n_gpus > 1 {
        check(unsafe { plonkish_cuda_bases_cached_sharded(n_gpus, bases.as_ptr() as *const c_void, bases.len(), &mut h) }, "plonkish_cuda_bases_cached_sharded");
    } else {
        check(unsafe { plonkish_cuda_bases_cached(0, bases.as_ptr() as *const c_void, bases.len(), &mut h) }, "plonkish_cuda_bases_cached");
    }
    h
}

/// Call from the Drop of whatever owns a slice that went through `variable_base_msm` (e.g. a ProverParam wrapper):
/// frees its resident table at once instead of waiting for the LRU bound.
pub fn forget_bases(bases: &[G1Affine]) {
    unsafe { plonkish_cuda_bases_cache_evict(bases.as_ptr() as *const c_void) };
}

/// An explicitly registered base slice: the handle lives exactly as long as this value (a field of a ProverParam
/// wrapper next to `eqs[k]` / `powers_of_s_g1`).  No address is ever used as a key.
pub struct RegisteredBases {
    handle: u64,
    len: usize,
    sharded: c_int,
}
impl RegisteredBases {
    pub fn new(bases: &[G1Affine]) -> Self {
        init();
        let mut handle = 0u64;
        let g = if bases.len() >= MULTI_MIN { gpus() } else { 1 };
        if g > 1 {
            check(unsafe { plonkish_cuda_bases_register_sharded(g, bases.as_ptr() as *const c_void, bases.len(), &mut handle) }, "plonkish_cuda_bases_register_sharded");
        } else {
            check(unsafe { plonkish_cuda_bases_register(0, bases.as_ptr() as *const c_void, bases.len(), &mut handle) }, "plonkish_cuda_bases_register");
        }
        Self { handle, len: bases.len(), sharded: g }
    }
    /// `variable_base_msm(scalars, &bases[..scalars.len()])` (kzg.rs:255, univariate/kzg.rs:28).
    pub fn msm(&self, scalars: &[Fr]) -> G1Affine {
        assert!(scalars.len() <= self.len); // msm.rs:90
        let mut out = [0u8; 64];
        if self.sharded > 1 {
            assert_eq!(scalars.len(), self.len, "a sharded slice serves its full length only");
            check(unsafe { plonkish_cuda_msm_bn254_g1_multi(self.sharded, scalars.as_ptr() as *const c_void, std::ptr::null(), self.handle, scalars.len(), out.as_mut_ptr() as *mut c_void) },
                  "plonkish_cuda_msm_bn254_g1_multi");
        } else {
            check(unsafe { plonkish_cuda_msm_bn254_g1(scalars.as_ptr() as *const c_void, std::ptr::null(), self.handle, scalars.len(), out.as_mut_ptr() as *mut c_void) },
                  "plonkish_cuda_msm_bn254_g1");
        }
        unsafe { std::mem::transmute(out) }
    }
}
impl Drop for RegisteredBases {
    fn drop(&mut self) {
        // reference counted inside the library: a call still in flight on another rayon worker finishes first
        unsafe { plonkish_cuda_bases_release(self.handle) };
    }
}

/// `variable_base_msm` for BN254 G1 (msm.rs:84-115).  Accepts the same iterators of
/// references; contiguous inputs (the commit paths, pcs/multilinear/kzg.rs:255,271,292;
/// pcs/univariate/kzg.rs:28) go down as two slices, anything else is gathered.  With
/// PLONKISH_CUDA_GPUS=G > 1, contiguous MSMs of 2^21 points and more are point-sharded over G GPUs
/// (one host thread per device, NCCL gather of the partials — plonkish_cuda_msm_bn254_g1_multi).
pub fn variable_base_msm_bn254<'a, 'b>(
    scalars: impl IntoIterator<Item = &'a Fr>,
    bases: impl IntoIterator<Item = &'b G1Affine>,
) -> G1 {
    init();
    let scalars: Vec<&Fr> = scalars.into_iter().collect();
    let bases: Vec<&G1Affine> = bases.into_iter().collect();
    assert_eq!(scalars.len(), bases.len()); // msm.rs:90
    let n = scalars.len();
    let mut out = [0u8; 64];
    if n == 0 {
        return G1::default(); // documented deviation: the reference panics at msm.rs:154
    }
    let contiguous = |first: usize, last: usize, stride: usize| last == first + (n - 1) * stride;
    let s0 = scalars[0] as *const Fr as usize;
    let b0 = bases[0] as *const G1Affine as usize;
    let s_contig = contiguous(s0, scalars[n - 1] as *const Fr as usize, 32)
        && scalars.windows(2).all(|w| (w[1] as *const Fr as usize) == (w[0] as *const Fr as usize) + 32);
    let b_contig = contiguous(b0, bases[n - 1] as *const G1Affine as usize, 64)
        && bases.windows(2).all(|w| (w[1] as *const G1Affine as usize) == (w[0] as *const G1Affine as usize) + 64);
    if s_contig && b_contig {
        // SAFETY: n consecutive G1Affine starting at bases[0], borrowed for the duration of this call.
        let base_slice = unsafe { std::slice::from_raw_parts(b0 as *const G1Affine, n) };
        let g = if n >= MULTI_MIN { gpus() } else { 1 };
        let handle = if n >= REGISTER_MIN { cached_handle(base_slice, g) } else { 0 };
        let bases_ptr = if handle == 0 { b0 as *const c_void } else { std::ptr::null() };
        if g > 1 {
            check(
                unsafe { plonkish_cuda_msm_bn254_g1_multi(g, s0 as *const c_void, bases_ptr, handle, n, out.as_mut_ptr() as *mut c_void) },
                "plonkish_cuda_msm_bn254_g1_multi",
            );
        } else {
            check(
                unsafe { plonkish_cuda_msm_bn254_g1(s0 as *const c_void, bases_ptr, handle, n, out.as_mut_ptr() as *mut c_void) },
                "plonkish_cuda_msm_bn254_g1",
            );
        }
    } else {
        let sp: Vec<*const c_void> = scalars.iter().map(|s| *s as *const Fr as *const c_void).collect();
        let bp: Vec<*const c_void> = bases.iter().map(|b| *b as *const G1Affine as *const c_void).collect();
        check(
            unsafe { plonkish_cuda_msm_bn254_g1_gather(sp.as_ptr(), bp.as_ptr(), n, out.as_mut_ptr() as *mut c_void) },
            "plonkish_cuda_msm_bn254_g1_gather",
        );
    }
    // The library returns the normalised point; every hot-path caller normalises anyway
    // (`.into()` kzg.rs:255,271,292; `.to_affine()` pcs.rs:175), so G1 is rebuilt from it.
    let affine: G1Affine = unsafe { std::mem::transmute(out) };
    affine.to_curve()
}

/// The loop of `MultilinearKzg::batch_commit` (pcs/multilinear/kzg.rs:259-274) in one call: every
/// polynomial has `n` evaluations and is committed against the same `bases` slice; the upload of
/// polynomial j+1 overlaps the MSM of polynomial j.  Returns the commitments in order.  With the `timer`
/// feature the library prints one `variable_base_msm-{n}` Start/End pair per polynomial (msm.rs:92).
pub fn batch_commit_bn254(polys: &[&[Fr]], bases: &[G1Affine]) -> Vec<G1Affine> {
    init();
    if polys.is_empty() {
        return Vec::new();
    }
    let n = polys[0].len();
    assert!(polys.iter().all(|p| p.len() == n) && n <= bases.len());
    let handle = cached_handle(bases, 1);
    let ptrs: Vec<*const c_void> = polys.iter().map(|p| p.as_ptr() as *const c_void).collect();
    let mut out = vec![[0u8; 64]; polys.len()];
    check(
        unsafe { plonkish_cuda_msm_bn254_g1_batch(ptrs.as_ptr(), polys.len(), handle, n, out.as_mut_ptr() as *mut c_void) },
        "plonkish_cuda_msm_bn254_g1_batch",
    );
    out.into_iter().map(|b| unsafe { std::mem::transmute::<[u8; 64], G1Affine>(b) }).collect()
}

/// The quotient commitments of `MultilinearKzg::open` (pcs/multilinear/kzg.rs:291-293) in one call:
/// `quotients[i]` (2^i scalars) is committed against `eqs[i]`.  The quotients do not depend on the
/// commitments, so `open` collects them first (pcs/multilinear.rs:72-107 with a collecting closure)
/// and commits afterwards; small MSMs run concurrently on the GPU.
pub fn commit_quotients_bn254(quotients: &[Vec<Fr>], eqs: &[Vec<G1Affine>]) -> Vec<G1Affine> {
    init();
    assert!(quotients.len() <= eqs.len());
    let handles: Vec<u64> = quotients.iter().zip(eqs).map(|(q, e)| {
        assert!(q.len() <= e.len());
        cached_handle(e, 1)
    }).collect();
    let ptrs: Vec<*const c_void> = quotients.iter().map(|q| q.as_ptr() as *const c_void).collect();
    let ns: Vec<usize> = quotients.iter().map(|q| q.len()).collect();
    let mut out = vec![[0u8; 64]; quotients.len()];
    check(
        unsafe { plonkish_cuda_msm_bn254_g1_many(ptrs.as_ptr(), handles.as_ptr(), ns.as_ptr(), quotients.len(), out.as_mut_ptr() as *mut c_void) },
        "plonkish_cuda_msm_bn254_g1_many",
    );
    out.into_iter().map(|b| unsafe { std::mem::transmute::<[u8; 64], G1Affine>(b) }).collect()
}

// ---- the callers either side of the MSM ------------------------

/// A polynomial's evaluations kept in HBM between `batch_commit` and `open`
/// (pcs/multilinear/kzg.rs:259-274 -> :276-302): what `poly.evals()` is to the reference.
pub struct ResidentPoly {
    handle: u64,
    num_vars: usize,
}
impl Drop for ResidentPoly {
    fn drop(&mut self) {
        unsafe { plonkish_cuda_scalars_release(self.handle) };
    }
}

/// The prover half of `MultilinearKzg::setup` (kzg.rs:167-212) built on the GPU: `eqs[k]` for
/// k = 0..=num_vars, registered as resident base slices and never materialised on the host.
pub struct ResidentEqs {
    handles: Vec<u64>,
}
impl ResidentEqs {
    pub fn setup(g1: &G1Affine, ss: &[Fr]) -> Self {
        init();
        let mut handles = vec![0u64; ss.len() + 1];
        check(
            unsafe { plonkish_cuda_kzg_setup_eqs_bn254(0, g1 as *const G1Affine as *const c_void, ss.as_ptr() as *const c_void, ss.len(), handles.as_mut_ptr()) },
            "plonkish_cuda_kzg_setup_eqs_bn254",
        );
        Self { handles }
    }
    pub fn num_vars(&self) -> usize {
        self.handles.len() - 1 // kzg.rs:68-70
    }
    /// `eqs[k]` back on the host, for `MultilinearKzgProverParams { g1, eqs }` serialisation (kzg.rs:55-77).
    pub fn eq_to_host(&self, k: usize) -> Vec<G1Affine> {
        let mut out = vec![G1Affine::default(); 1 << k];
        check(unsafe { plonkish_cuda_bases_read(self.handles[k], 0, out.len(), out.as_mut_ptr() as *mut c_void) }, "plonkish_cuda_bases_read");
        out
    }
    /// `batch_commit` (kzg.rs:259-274) of polynomials with `num_vars` variables each, leaving them resident.
    pub fn batch_commit_keep(&self, polys: &[&[Fr]]) -> (Vec<G1Affine>, Vec<ResidentPoly>) {
        let n = polys[0].len();
        assert!(n.is_power_of_two() && polys.iter().all(|p| p.len() == n));
        let k = n.trailing_zeros() as usize;
        let ptrs: Vec<*const c_void> = polys.iter().map(|p| p.as_ptr() as *const c_void).collect();
        let mut out = vec![G1Affine::default(); polys.len()];
        let mut hs = vec![0u64; polys.len()];
        check(
            unsafe { plonkish_cuda_msm_bn254_g1_batch_keep(ptrs.as_ptr(), polys.len(), self.handles[k], n, out.as_mut_ptr() as *mut c_void, hs.as_mut_ptr()) },
            "plonkish_cuda_msm_bn254_g1_batch_keep",
        );
        (out, hs.into_iter().map(|handle| ResidentPoly { handle, num_vars: k }).collect())
    }
    /// `MultilinearKzg::open` (kzg.rs:276-302): the quotient commitments in transcript order and f(point).
    pub fn open(&self, poly: &ResidentPoly, point: &[Fr]) -> (Vec<G1Affine>, Fr) {
        assert_eq!(poly.num_vars, point.len()); // multilinear.rs:77
        let mut comms = vec![G1Affine::default(); point.len()];
        let mut eval = Fr::zero();
        check(
            unsafe {
                plonkish_cuda_kzg_open_bn254(poly.handle, self.handles.as_ptr(), point.as_ptr() as *const c_void, point.len(),
                                             comms.as_mut_ptr() as *mut c_void, &mut eval as *mut Fr as *mut c_void)
            },
            "plonkish_cuda_kzg_open_bn254",
        );
        (comms, eval)
    }
}
impl Drop for ResidentEqs {
    fn drop(&mut self) {
        for h in &self.handles {
            unsafe { plonkish_cuda_bases_release(*h) };
        }
    }
}

/// The G1 half of `UnivariateKzg::setup` (pcs/univariate/kzg.rs:175-195): `powers_of_s_g1`, built on the GPU and read back
/// for `UnivariateKzgParam` (`commit_coeffs` on the returned vector goes through the content-validated cache like any
/// other borrowed slice).
pub fn univariate_powers_of_s_g1(g1: &G1Affine, s: &Fr, poly_size: usize) -> Vec<G1Affine> {
    init();
    let mut handle = 0u64;
    check(
        unsafe { plonkish_cuda_kzg_setup_powers_bn254(0, g1 as *const G1Affine as *const c_void, s as *const Fr as *const c_void, poly_size, &mut handle) },
        "plonkish_cuda_kzg_setup_powers_bn254",
    );
    let mut out = vec![G1Affine::default(); poly_size];
    check(unsafe { plonkish_cuda_bases_read(handle, 0, poly_size, out.as_mut_ptr() as *mut c_void) }, "plonkish_cuda_bases_read");
    unsafe { plonkish_cuda_bases_release(handle) };
    out
}

/// `MultilinearPolynomial::eq_xy(y)` as a resident table (the zero-check factor of piop/sum_check/classic.rs:57-61).
pub fn eq_xy(y: &[Fr]) -> ResidentPoly {
    init();
    let mut handle = 0u64;
    check(unsafe { plonkish_cuda_eq_table(0, y.as_ptr() as *const c_void, y.len(), &mut handle) }, "plonkish_cuda_eq_table");
    ResidentPoly { handle, num_vars: y.len() }
}

/// The g_prime merge of `batch_open` (pcs/multilinear.rs:203-213): sum_i coeffs[i] * polys[i], in HBM.
pub fn linear_combination(polys: &[&ResidentPoly], coeffs: &[Fr]) -> ResidentPoly {
    assert_eq!(polys.len(), coeffs.len());
    let hs: Vec<u64> = polys.iter().map(|p| p.handle).collect();
    let num_vars = polys[0].num_vars;
    let mut handle = 0u64;
    check(
        unsafe { plonkish_cuda_fr_linear_combination(hs.as_ptr(), coeffs.as_ptr() as *const c_void, hs.len(), 1 << num_vars, &mut handle) },
        "plonkish_cuda_fr_linear_combination",
    );
    ResidentPoly { handle, num_vars }
}

/// `permutation_z_polys` (backend/hyperplonk/prover.rs:252-345) on resident polynomials: `values[i]` is the witness column
/// `polys[*poly]` of permutation polynomial i, `sigmas[i]` the permutation polynomial itself (in `pp.permutation_polys`
/// order); returns `num_chunks` resident z polynomials, ready for `batch_commit` (hyperplonk.rs:251-252).
pub fn permutation_z_polys(num_chunks: usize, values: &[&ResidentPoly], sigmas: &[&ResidentPoly], beta: &Fr, gamma: &Fr) -> Vec<ResidentPoly> {
    assert_eq!(values.len(), sigmas.len());
    let num_vars = values[0].num_vars;
    let vh: Vec<u64> = values.iter().map(|p| p.handle).collect();
    let sh: Vec<u64> = sigmas.iter().map(|p| p.handle).collect();
    let mut out = vec![0u64; num_chunks];
    check(
        unsafe {
            plonkish_cuda_permutation_z_polys_bn254(vh.as_ptr(), sh.as_ptr(), vh.len(), num_chunks, num_vars, beta as *const Fr as *const c_void,
                                                    gamma as *const Fr as *const c_void, out.as_mut_ptr())
        },
        "plonkish_cuda_permutation_z_polys_bn254",
    );
    out.into_iter().map(|handle| ResidentPoly { handle, num_vars }).collect()
}

/// One table of a compiled zero-check expression: `constant + identity_coeff * id + sum_i coeffs[i] * polys[i]` rotated by
/// `rotations[i]` (BooleanHypercube::rotate, util/arithmetic/bh.rs:104-121), plus `values[j]` on row `rows[j]` (a Lagrange
/// polynomial is one such row, an instance polynomial a handful: backend/hyperplonk/prover.rs:32-48).  What
/// `ProverState` keeps implicit (piop/sum_check/classic.rs:40-75, 104-126) as an explicit multilinear table; the linear
/// factors `w + beta * id + gamma` of the permutation constraint (preprocessor.rs:153-165) are one call each.
pub fn affine_table(num_vars: usize, polys: &[&ResidentPoly], coeffs: &[Fr], rotations: &[i32], constant: Option<&Fr>, identity_coeff: Option<&Fr>,
                    rows: &[u64], values: &[Fr]) -> ResidentPoly {
    init();
    assert!(polys.len() == coeffs.len() && polys.len() == rotations.len() && rows.len() == values.len());
    let hs: Vec<u64> = polys.iter().map(|p| p.handle).collect();
    let as_ptr = |f: Option<&Fr>| f.map_or(std::ptr::null(), |f| f as *const Fr as *const c_void);
    let mut handle = 0u64;
    check(
        unsafe {
            plonkish_cuda_fr_affine_table(0, num_vars, hs.as_ptr(), rotations.as_ptr(), coeffs.as_ptr() as *const c_void, hs.len(), as_ptr(constant),
                                          as_ptr(identity_coeff), rows.as_ptr(), values.as_ptr() as *const c_void, rows.len(), &mut handle)
        },
        "plonkish_cuda_fr_affine_table",
    );
    ResidentPoly { handle, num_vars }
}

/// A compiled expression on every row — `lookup_compressed_poly` (backend/hyperplonk/prover.rs:79-137) is one call:
/// `terms[t] = (coeff, indices into polys)`, out[b] = sum_t coeff_t * prod_j polys[idx][b] (times `polys[common][b]`).
pub fn expression_table(polys: &[&ResidentPoly], terms: &[(Fr, Vec<u32>)], common: Option<usize>) -> ResidentPoly {
    let hs: Vec<u64> = polys.iter().map(|p| p.handle).collect();
    let coeffs: Vec<Fr> = terms.iter().map(|(c, _)| *c).collect();
    let mut offsets = vec![0u32];
    let mut flat = Vec::new();
    for (_, idx) in terms {
        flat.extend_from_slice(idx);
        offsets.push(flat.len() as u32);
    }
    let num_vars = polys[0].num_vars;
    let mut handle = 0u64;
    check(
        unsafe {
            plonkish_cuda_fr_expression_table(hs.as_ptr(), hs.len(), num_vars, coeffs.as_ptr() as *const c_void, offsets.as_ptr(), flat.as_ptr(), terms.len(),
                                              common.map_or(-1, |c| c as c_int), &mut handle)
        },
        "plonkish_cuda_fr_expression_table",
    );
    ResidentPoly { handle, num_vars }
}

/// `lookup_m_poly` (backend/hyperplonk/prover.rs:145-192); `Err` when an input value is not in the table, like the
/// reference's `Error::InvalidSnark("Invalid lookup input")`.
pub fn lookup_m_poly(input: &ResidentPoly, table: &ResidentPoly) -> Result<ResidentPoly, String> {
    let mut handle = 0u64;
    let rc = unsafe { plonkish_cuda_lookup_m_poly_bn254(input.handle, table.handle, &mut handle) };
    if rc != 0 {
        return Err(unsafe { std::ffi::CStr::from_ptr(plonkish_cuda_last_error()) }.to_string_lossy().into_owned());
    }
    Ok(ResidentPoly { handle, num_vars: input.num_vars })
}

/// `lookup_h_poly` (backend/hyperplonk/prover.rs:206-250): 1 / (gamma + input) - m / (gamma + table).
pub fn lookup_h_poly(input: &ResidentPoly, table: &ResidentPoly, m: &ResidentPoly, gamma: &Fr) -> ResidentPoly {
    let mut handle = 0u64;
    check(
        unsafe { plonkish_cuda_lookup_h_poly_bn254(input.handle, table.handle, m.handle, gamma as *const Fr as *const c_void, &mut handle) },
        "plonkish_cuda_lookup_h_poly_bn254",
    );
    ResidentPoly { handle, num_vars: input.num_vars }
}

/// `MultilinearPolynomial::evaluate` at several points (poly/multilinear.rs:137-156), e.g. the 2^distance points of
/// `evaluate_for_rotation` (poly/multilinear.rs:191-263) that `prove_sum_check` writes to the transcript (prover.rs:392-406).
pub fn evaluate(poly: &ResidentPoly, points: &[Vec<Fr>]) -> Vec<Fr> {
    let flat: Vec<Fr> = points.iter().flat_map(|p| { assert_eq!(p.len(), poly.num_vars); p.iter().copied() }).collect();
    let mut out = vec![Fr::zero(); points.len()];
    check(
        unsafe { plonkish_cuda_fr_evaluate(poly.handle, flat.as_ptr() as *const c_void, poly.num_vars, points.len(), out.as_mut_ptr() as *mut c_void) },
        "plonkish_cuda_fr_evaluate",
    );
    out
}

/// Coefficients of a univariate polynomial kept in HBM, for `UnivariateKzg::open` / `batch_open` (pcs/univariate/kzg.rs:264-354).
pub struct ResidentCoeffs {
    handle: u64,
    len: usize,
}
impl Drop for ResidentCoeffs {
    fn drop(&mut self) {
        unsafe { plonkish_cuda_scalars_release(self.handle) };
    }
}
impl ResidentCoeffs {
    pub fn new(coeffs: &[Fr]) -> Self {
        init();
        let mut handle = 0u64;
        check(unsafe { plonkish_cuda_scalars_register(0, coeffs.as_ptr() as *const c_void, coeffs.len(), &mut handle) }, "plonkish_cuda_scalars_register");
        Self { handle, len: coeffs.len() }
    }
    /// `poly.div_rem(&(X - z))` (poly/univariate.rs:144-168): the quotient stays resident (same length, zero top coefficient).
    pub fn div_linear(&self, z: &Fr) -> (ResidentCoeffs, Fr) {
        let (mut q, mut rem) = (0u64, Fr::zero());
        check(
            unsafe { plonkish_cuda_fr_div_linear(self.handle, z as *const Fr as *const c_void, &mut q, &mut rem as *mut Fr as *mut c_void) },
            "plonkish_cuda_fr_div_linear",
        );
        (ResidentCoeffs { handle: q, len: self.len }, rem)
    }
    /// `commit_coeffs` (pcs/univariate/kzg.rs:24-30) against a registered `powers_of_s_g1`.
    pub fn commit(&self, powers_of_s_g1: &RegisteredBases) -> G1Affine {
        assert!(self.len <= powers_of_s_g1.len);
        let mut out = [0u8; 64];
        check(
            unsafe { plonkish_cuda_msm_bn254_g1_resident(self.handle, powers_of_s_g1.handle, self.len, out.as_mut_ptr() as *mut c_void) },
            "plonkish_cuda_msm_bn254_g1_resident",
        );
        unsafe { std::mem::transmute(out) }
    }
    /// `UnivariateKzg::open` (kzg.rs:264-299): the quotient's commitment (what `write_commitment` takes) and poly(z).
    pub fn open(&self, powers_of_s_g1: &RegisteredBases, z: &Fr) -> (G1Affine, Fr) {
        let (quotient, eval) = self.div_linear(z);
        (quotient.commit(powers_of_s_g1), eval)
    }
}

/// The polynomial work of `Zeromorph::<UnivariateKzg<Bn256>>::open` (pcs/multilinear/zeromorph.rs:126-186) on a resident
/// polynomial; the caller keeps the transcript and the scalar work (`eval_and_quotient_scalars`, :259-294).
pub struct ZeromorphQuotients {
    handle: u64, // packed: quotient i (2^i values) at element offset 2^i
    num_vars: usize,
}
impl Drop for ZeromorphQuotients {
    fn drop(&mut self) {
        unsafe { plonkish_cuda_scalars_release(self.handle) };
    }
}
impl ZeromorphQuotients {
    /// `quotients(poly, point, ..)` (pcs/multilinear.rs:72-107; zeromorph.rs:149): the quotients stay in HBM, the remainder comes back.
    pub fn new(poly: &ResidentPoly, point: &[Fr]) -> (Self, Fr) {
        assert_eq!(poly.num_vars, point.len());
        let (mut handle, mut eval) = (0u64, Fr::zero());
        check(
            unsafe { plonkish_cuda_fr_quotients(poly.handle, point.as_ptr() as *const c_void, point.len(), &mut handle, &mut eval as *mut Fr as *mut c_void) },
            "plonkish_cuda_fr_quotients",
        );
        (Self { handle, num_vars: point.len() }, eval)
    }
    /// `UnivariateKzg::batch_commit` of the quotients (zeromorph.rs:150): q_i against `powers_of_s_g1[..2^i]`, one call.
    pub fn commit(&self, powers_of_s_g1: &RegisteredBases) -> Vec<G1Affine> {
        let sizes: Vec<usize> = (0..self.num_vars).map(|i| 1usize << i).collect();
        let handles = vec![powers_of_s_g1.handle; self.num_vars];
        let mut out = vec![G1Affine::default(); self.num_vars];
        check(
            unsafe {
                plonkish_cuda_msm_bn254_g1_many_resident(self.handle, sizes.as_ptr(), handles.as_ptr(), sizes.as_ptr(), self.num_vars, out.as_mut_ptr() as *mut c_void)
            },
            "plonkish_cuda_msm_bn254_g1_many_resident",
        );
        out
    }
    /// `q_hat` (zeromorph.rs:157-168) for the challenge powers `powers(y).take(num_vars)`.
    pub fn q_hat(&self, powers_of_y: &[Fr]) -> ResidentCoeffs {
        assert_eq!(powers_of_y.len(), self.num_vars);
        let mut handle = 0u64;
        check(
            unsafe { plonkish_cuda_zeromorph_q_hat_bn254(self.handle, powers_of_y.as_ptr() as *const c_void, self.num_vars, &mut handle) },
            "plonkish_cuda_zeromorph_q_hat_bn254",
        );
        ResidentCoeffs { handle, len: 1 << self.num_vars }
    }
    /// `f` (zeromorph.rs:175-180): z * poly + q_hat, f[0] += c0 (= eval_scalar * eval), f += (q_scalars[i], q_i).
    pub fn f(&self, poly: &ResidentPoly, q_hat: &ResidentCoeffs, z: &Fr, c0: &Fr, q_scalars: &[Fr]) -> ResidentCoeffs {
        assert_eq!(q_scalars.len(), self.num_vars);
        let mut handle = 0u64;
        check(
            unsafe {
                plonkish_cuda_zeromorph_f_bn254(poly.handle, q_hat.handle, self.handle, z as *const Fr as *const c_void, c0 as *const Fr as *const c_void,
                                                q_scalars.as_ptr() as *const c_void, self.num_vars, &mut handle)
            },
            "plonkish_cuda_zeromorph_f_bn254",
        );
        ResidentCoeffs { handle, len: 1 << self.num_vars }
    }
}

/// The folds of `Gemini::<UnivariateKzg<Bn256>>::open` (pcs/multilinear/gemini.rs:98-108) kept in HBM: `fs[1..]`, packed.
pub struct GeminiFolds {
    handle: u64, // f_i (2^(num_vars - i) coefficients) at element offset 2^(num_vars - i)
    num_vars: usize,
}
impl Drop for GeminiFolds {
    fn drop(&mut self) {
        unsafe { plonkish_cuda_scalars_release(self.handle) };
    }
}
impl GeminiFolds {
    pub fn new(poly: &ResidentPoly, point: &[Fr]) -> Self {
        assert!(poly.num_vars == point.len() && !point.is_empty());
        let mut handle = 0u64;
        check(unsafe { plonkish_cuda_fr_gemini_folds(poly.handle, point.as_ptr() as *const c_void, point.len(), &mut handle) }, "plonkish_cuda_fr_gemini_folds");
        Self { handle, num_vars: point.len() }
    }
    /// `UnivariateKzg::batch_commit(pp, &fs[1..])` (gemini.rs:124-128): f_i against `powers_of_s_g1[..2^(num_vars - i)]`, one call.
    pub fn commit(&self, powers_of_s_g1: &RegisteredBases) -> Vec<G1Affine> {
        let sizes: Vec<usize> = (1..self.num_vars).map(|i| 1usize << (self.num_vars - i)).collect();
        let handles = vec![powers_of_s_g1.handle; sizes.len()];
        let mut out = vec![G1Affine::default(); sizes.len()];
        if sizes.is_empty() {
            return out;
        }
        check(
            unsafe { plonkish_cuda_msm_bn254_g1_many_resident(self.handle, sizes.as_ptr(), handles.as_ptr(), sizes.as_ptr(), sizes.len(), out.as_mut_ptr() as *mut c_void) },
            "plonkish_cuda_msm_bn254_g1_many_resident",
        );
        out
    }
    /// `fs[i]` (1 <= i < num_vars) as coefficients of its own for `UnivariateKzg::batch_open` (gemini.rs:140); shares the memory.
    pub fn fold(&self, i: usize) -> ResidentCoeffs {
        assert!(i >= 1 && i < self.num_vars);
        let len = 1usize << (self.num_vars - i);
        let mut handle = 0u64;
        check(unsafe { plonkish_cuda_scalars_slice(self.handle, len, len, &mut handle) }, "plonkish_cuda_scalars_slice");
        ResidentCoeffs { handle, len }
    }
}

/// `sum_i coeffs[i] * polys[i]` over coefficient vectors of different lengths (`f += (scalar, q)`, poly/univariate.rs):
/// the sums of `UnivariateKzg::batch_open` (pcs/univariate/kzg.rs:330, 339-343) over Gemini's folds.
pub fn linear_combination_padded(polys: &[&ResidentCoeffs], coeffs: &[Fr]) -> ResidentCoeffs {
    assert!(!polys.is_empty() && polys.len() == coeffs.len());
    let handles: Vec<u64> = polys.iter().map(|p| p.handle).collect();
    let len = polys.iter().map(|p| p.len).max().unwrap();
    let mut handle = 0u64;
    check(
        unsafe { plonkish_cuda_fr_linear_combination_padded(handles.as_ptr(), coeffs.as_ptr() as *const c_void, polys.len(), len, &mut handle) },
        "plonkish_cuda_fr_linear_combination_padded",
    );
    ResidentCoeffs { handle, len }
}

/// `fixed_base_msm(window_size, &window_table(window_size, base), scalars)` followed by
/// `batch_normalize` (msm.rs:16-31, 67-81; kzg.rs:195-208; univariate/kzg.rs:187-199).
pub fn fixed_base_msm_bn254(base: &G1Affine, scalars: &[Fr]) -> Vec<G1Affine> {
    init();
    let mut out = vec![G1Affine::default(); scalars.len()];
    check(
        unsafe {
            plonkish_cuda_fixed_base_msm_bn254_g1(0, base as *const G1Affine as *const c_void, scalars.as_ptr() as *const c_void, scalars.len(),
                                                  out.as_mut_ptr() as *mut c_void)
        },
        "plonkish_cuda_fixed_base_msm_bn254_g1",
    );
    out
}

/// `ClassicSumCheck<EvaluationsProver>::prove` (piop/sum_check/classic.rs:208-240) over resident tables.
/// `terms[t] = (coeff, factors)` is the flattened expression sum_t coeff * prod polys[factor]; `common` is the
/// table multiplying the whole sum (the eq(x, y) of a zero check).  `squeeze` stands for
/// `msg.write(transcript)?; transcript.squeeze_challenge()` (classic.rs:226-229) and `evaluate` for
/// `msg.evaluate(&aux, &challenge)` (eval.rs:50-52): both stay with the caller's transcript and field code.
pub fn sum_check_prove(
    polys: &[&ResidentPoly],
    terms: &[(Fr, Vec<u32>)],
    common: Option<usize>,
    sum: Fr,
    mut squeeze: impl FnMut(&[Fr]) -> Fr,
    evaluate: impl Fn(&[Fr], &Fr) -> Fr,
) -> (Vec<Fr>, Vec<Fr>) {
    let num_vars = polys[0].num_vars;
    let hs: Vec<u64> = polys.iter().map(|p| p.handle).collect();
    let coeffs: Vec<Fr> = terms.iter().map(|t| t.0).collect();
    let mut offsets = vec![0u32];
    let mut flat = Vec::new();
    for (_, f) in terms {
        flat.extend_from_slice(f);
        offsets.push(flat.len() as u32);
    }
    let mut state = 0u64;
    check(
        unsafe {
            plonkish_cuda_sumcheck_new(hs.as_ptr(), hs.len(), num_vars, coeffs.as_ptr() as *const c_void, offsets.as_ptr(), flat.as_ptr(), terms.len(),
                                       common.map_or(-1, |c| c as c_int), &mut state)
        },
        "plonkish_cuda_sumcheck_new",
    );
    let degree = unsafe { plonkish_cuda_sumcheck_degree(state) } as usize;
    let (mut sum, mut challenges) = (sum, Vec::with_capacity(num_vars));
    for _ in 0..num_vars {
        let mut msg = vec![Fr::zero(); degree + 1];
        check(unsafe { plonkish_cuda_sumcheck_round(state, msg[1..].as_mut_ptr() as *mut c_void) }, "plonkish_cuda_sumcheck_round");
        msg[0] = sum - msg[1]; // eval.rs:128
        let challenge = squeeze(&msg);
        sum = evaluate(&msg, &challenge); // classic.rs:232
        check(unsafe { plonkish_cuda_sumcheck_fix_var(state, &challenge as *const Fr as *const c_void) }, "plonkish_cuda_sumcheck_fix_var");
        challenges.push(challenge);
    }
    let mut evals = vec![Fr::zero(); polys.len()];
    check(unsafe { plonkish_cuda_sumcheck_final_evals(state, evals.as_mut_ptr() as *mut c_void) }, "plonkish_cuda_sumcheck_final_evals");
    unsafe { plonkish_cuda_sumcheck_free(state) };
    (challenges, evals) // classic.rs:239
}

/// The zero check of `prove_zero_check` (backend/hyperplonk/prover.rs:347-365): `sum_check_prove` for an expression whose
/// common table `polys[common]` is `eq_xy(y)` (classic.rs:57-61), run factored.  In round r the common factor's pair is
/// (S (1 - y_r), S y_r), so the round polynomial is h(X) = (1 - y_r + X (2 y_r - 1)) G(X); the GPU returns G(1..degree-1)
/// (`plonkish_cuda_sumcheck_round_factored`: one evaluation point fewer per pair), G(0) follows from h(0) + h(1) = sum and
/// G(degree) from `evaluate` on G's own points.  The messages written are the ones `sum_check_prove` writes.
pub fn zero_check_prove(
    polys: &[&ResidentPoly],
    terms: &[(Fr, Vec<u32>)],
    common: usize,
    y: &[Fr],
    mut squeeze: impl FnMut(&[Fr]) -> Fr,
    evaluate: impl Fn(&[Fr], &Fr) -> Fr,
) -> (Vec<Fr>, Vec<Fr>) {
    use halo2_curves::ff::Field;
    let num_vars = polys[0].num_vars;
    assert_eq!(y.len(), num_vars);
    let hs: Vec<u64> = polys.iter().map(|p| p.handle).collect();
    let coeffs: Vec<Fr> = terms.iter().map(|t| t.0).collect();
    let mut offsets = vec![0u32];
    let mut flat = Vec::new();
    for (_, f) in terms {
        flat.extend_from_slice(f);
        offsets.push(flat.len() as u32);
    }
    let mut state = 0u64;
    check(
        unsafe {
            plonkish_cuda_sumcheck_new(hs.as_ptr(), hs.len(), num_vars, coeffs.as_ptr() as *const c_void, offsets.as_ptr(), flat.as_ptr(), terms.len(),
                                       common as c_int, &mut state)
        },
        "plonkish_cuda_sumcheck_new",
    );
    let degree = unsafe { plonkish_cuda_sumcheck_degree(state) } as usize;
    let (mut sum, mut challenges) = (Fr::zero(), Vec::with_capacity(num_vars));
    for y_r in y {
        let (e0, e1) = (Fr::one() - y_r, *y_r);
        let mut msg = vec![Fr::zero(); degree + 1];
        match Option::<Fr>::from(e0.invert()) {
            Some(inv_e0) if degree >= 2 => {
                let mut g = vec![Fr::zero(); degree]; // G(0..degree-1)
                check(unsafe { plonkish_cuda_sumcheck_round_factored(state, g[1..].as_mut_ptr() as *mut c_void) }, "plonkish_cuda_sumcheck_round_factored");
                g[0] = (sum - e1 * g[1]) * inv_e0;
                let g_top = evaluate(&g, &Fr::from(degree as u64));
                g.push(g_top);
                for (x, (m, gx)) in msg.iter_mut().zip(g.iter()).enumerate() {
                    *m = (e0 + Fr::from(x as u64) * (e1 - e0)) * gx;
                }
            }
            _ => {
                check(unsafe { plonkish_cuda_sumcheck_round(state, msg[1..].as_mut_ptr() as *mut c_void) }, "plonkish_cuda_sumcheck_round");
                msg[0] = sum - msg[1];
            }
        }
        let challenge = squeeze(&msg);
        sum = evaluate(&msg, &challenge);
        check(unsafe { plonkish_cuda_sumcheck_fix_var(state, &challenge as *const Fr as *const c_void) }, "plonkish_cuda_sumcheck_fix_var");
        challenges.push(challenge);
    }
    let mut evals = vec![Fr::zero(); polys.len()];
    check(unsafe { plonkish_cuda_sumcheck_final_evals(state, evals.as_mut_ptr() as *mut c_void) }, "plonkish_cuda_sumcheck_final_evals");
    unsafe { plonkish_cuda_sumcheck_free(state) };
    (challenges, evals)
}
