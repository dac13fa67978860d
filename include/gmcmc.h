/* gmcmc.h — C ABI of the B200-native many-chain sampling hot path of general-mcmc.
 *
 * The reference crate (SauersML/general-mcmc 0.8.0) has no FFI: its boundary is the Rust trait /
 * constructor surface.  A thin Rust shim keeps those signatures and binds the entry
 * points below; each entry point cites the reference interface it replaces (file:line into
 * /root/reference/src).  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Conventions
 *   - every function returns gmcmc_status; nothing throws or aborts across the boundary (the crate's
 *     release profile is panic = "abort", Cargo.toml:54); detail via gmcmc_last_error() (thread-local).
 *   - inputs are borrowed for the call and copied; outputs go to caller-owned host buffers, or to
 *     library-owned device buffers for the *_device calls (valid until the next run or the destroy).
 *   - a sampler handle is NOT thread-safe (mirrors `&mut self`); distinct handles may be used from
 *     distinct threads.  One context = one GPU = one process rank.
 *   - there is NO CPU fallback: every compute entry point fails with GMCMC_ERR_CUDA when no sm_100
 *     device is usable.
 *
 * Sample layout: [chains, samples, dim], C-contiguous (core.rs:219-229, batched_hmc.rs:98-110,
 * hmc.rs:179-180).  MH output is always f64 (Trace -> f64, core.rs:34-51); HMC/NUTS output is the
 * sampler's scalar type.
 *
 * RNG contract (replaces rand::SmallRng / rand_distr ziggurat, whose streams are third-party and
 * unpinned by any reference test — SURVEY §8c): Philox4x32-10, key = (seed_lo, seed_hi),
 *   counter = (gchain_lo, gchain_hi, transition_index, (stream << 24) | block)
 * where gchain = chain_offset + local chain index (so results do not depend on how chains are sharded
 * over GPUs) and transition_index counts transitions since the sampler was created / re-seeded.
 *   stream 0: N(0,1) draws for the momentum (HMC/NUTS) or the proposal noise (MH);
 *             f32: block b -> elements 4b..4b+3 (Box–Muller on (r0,r1) and (r2,r3));
 *             f64: block b -> elements 2b, 2b+1 (53-bit uniforms from (r0,r1) and (r2,r3));
 *   stream 1: block 0 -> accept uniform (r0 [f32] or (r0,r1) [f64]); NUTS: r2/r3 -> Exp(1) draw (-ln of the 53-bit
 *             uniform; f32 samplers in fast math mode: -ln of the 24-bit uniform of r2);
 *   stream 2: NUTS tree uniforms, block = draw index / 4 (see gmcmc_nuts_create).
 *   stream 3: NUTS momentum probe after a mass-matrix update (see gmcmc_nuts_set_mass_adaptation).
 * Uniforms are in (0,1]:  f32 ((r>>8)+1)*2^-24,  f64 ((r64>>11)+1)*2^-53.
 * Fast-mode MH with dim == 2 feeds TWO transitions from one block of stream 0 (Philox4x32-10 is 40 of the kernel's
 * ~90 instructions per step): with t the absolute transition index, the counter word "transition_index" is t >> 1,
 * and transition t takes words (0, 1) when t is even, words (2, 3) when odd.  Of its word pair (w0, w1): bits 31..12
 * of w0 -> radius uniform (k + 1/2) 2^-20, bits 31..12 of w1 -> angle 2 pi (k + 1/2) 2^-20 - pi (the grid maps onto
 * itself under z -> -z: the proposal stays exactly symmetric), ((w0 & 0xfff) << 11) | (w1 & 0x7ff) -> the accept
 * uniform (k + 1/2) 2^-23, which is the whole uniform of this mode.  Exact mode keeps the layout above.
 */
#ifndef GMCMC_H
#define GMCMC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmcmc_ctx gmcmc_ctx;
typedef struct gmcmc_target gmcmc_target;
typedef struct gmcmc_sampler gmcmc_sampler;

typedef enum {
  GMCMC_OK = 0,
  GMCMC_ERR_INVALID = 1,     /* bad argument / shape */
  GMCMC_ERR_CUDA = 2,        /* CUDA runtime error or no usable device */
  GMCMC_ERR_UNSUPPORTED = 3, /* combination not implemented (e.g. dim too large for a kernel) */
  GMCMC_ERR_NCCL = 4,
  GMCMC_ERR_STATE = 5        /* call not valid in the sampler's current state */
} gmcmc_status;

typedef enum { GMCMC_F32 = 0, GMCMC_F64 = 1 } gmcmc_dtype;

/* Built-in targets.  `params` of gmcmc_target_create are doubles, cast once to the sampler dtype
 * (as the reference constructors do). */
typedef enum {
  GMCMC_TARGET_ISO_GAUSS = 0,     /* distributions.rs:398-406  params: [std]                         any dim */
  GMCMC_TARGET_GAUSS2D = 1,       /* distributions.rs:191-208  params: [m0,m1, a,b,c,d] (cov rows)   dim 2   */
  GMCMC_TARGET_DIFF_GAUSS2D = 2,  /* distributions.rs:215-291  params: [m0,m1, c00,c01,c10,c11]      dim 2   */
  GMCMC_TARGET_DENSE_GAUSS = 3,   /* N-D form of :265-291      params: [mu[d], P[d*d] row-major, norm_const]  */
  GMCMC_TARGET_ROSENBROCK2D = 4,  /* distributions.rs:495-515  params: [a,b]                         dim 2   */
  GMCMC_TARGET_ROSENBROCK_ND = 5, /* distributions.rs:535-555  params: none                          any dim */
  GMCMC_TARGET_GAUSS_MIXTURE = 6  /* isotropic mixture         params: [K, sigma, w[K], mu[K*d]]     any dim */
} gmcmc_target_kind;

typedef enum {
  GMCMC_MATH_FAST = 0,  /* FMA contraction, merged half-kicks, shuffle-tree reductions */
  GMCMC_MATH_EXACT = 1  /* reference operation order, no contraction, sequential sums: bit-for-bit the
                           CPU oracle for transcendental-free targets (parity mode) */
} gmcmc_math_mode;

typedef enum {
  GMCMC_ADAPT_NONE = 0,      /* fixed step size: the reference's HMC (batched_hmc.rs:35) */
  GMCMC_ADAPT_PER_CHAIN = 1, /* per-chain dual averaging: the reference's NUTS (generic_nuts.rs:882-924) */
  GMCMC_ADAPT_POOLED = 2     /* one step size from the mean acceptance statistic over ALL chains of all
                                ranks (NCCL all-reduce); new mode named by BASELINE.json north_star */
} gmcmc_adapt_mode;

/* RunStats / BasicStats, stats.rs:370-415 (name field dropped: "ESS" / "Split R-hat"). */
typedef struct { float min, median, max, mean, std; } gmcmc_basic_stats;
typedef struct {
  gmcmc_basic_stats ess;
  gmcmc_basic_stats rhat;     /* reference orientation sqrt(W / var_hat), stats.rs:452-454 */
  gmcmc_basic_stats rhat_std; /* standard orientation sqrt(var_hat / W), reported alongside */
} gmcmc_run_stats_t;

typedef struct {
  uint64_t transitions;  /* chain-transitions taken by this rank since creation */
  uint64_t accepts;      /* accepted proposals (this rank); NUTS: transitions that accepted a subtree proposal */
  uint64_t grad_evals;   /* algorithmic gradient evaluations: HMC L per transition; NUTS leapfrogs taken */
  uint64_t divergences;  /* non-finite log-accept (HMC) / s' = false by energy (NUTS) */
  double step_size;      /* current (pooled or mean per-chain) step size */
  double kernel_ms;      /* device time of the sampling kernels of the last run (CUDA events) */
  uint64_t launches;     /* kernel launches issued by the last run */
} gmcmc_counters;

/* ---- context ------------------------------------------------------------------------------- */
/* One context per process/GPU.  rank/world describe the chain sharding: rank r of
 * `world` owns a contiguous global chain range given per sampler by chain_offset. */
gmcmc_status gmcmc_ctx_create(int device, gmcmc_ctx** out);
/* Multi-GPU: `nccl_id` is the 128-byte ncclUniqueId produced by gmcmc_nccl_unique_id on rank 0 and
 * distributed by the host (torch.distributed / MPI / the Rust shim). */
gmcmc_status gmcmc_nccl_unique_id(void* out128);
gmcmc_status gmcmc_ctx_create_dist(int device, int rank, int world, const void* nccl_id, gmcmc_ctx** out);
gmcmc_status gmcmc_ctx_destroy(gmcmc_ctx*);
gmcmc_status gmcmc_ctx_synchronize(gmcmc_ctx*);
/* the CUDA stream all work of this context is issued on (cudaStream_t as void*) */
gmcmc_status gmcmc_ctx_stream(gmcmc_ctx*, void** out_stream);

/* sum-all-reduce of a small host f64 vector over the ranks of a distributed context (no-op when
 * world == 1); used by the host shim to combine per-rank counters. */
gmcmc_status gmcmc_ctx_all_reduce_f64(gmcmc_ctx*, double* host_inout, size_t n);
/* page-locked host memory for sample tensors (so gmcmc_run copies at full PCIe rate) */
gmcmc_status gmcmc_host_alloc(size_t bytes, void** out);
gmcmc_status gmcmc_host_free(void* p);

/* measured FP32 FMA-pipe peak of this GPU in TFLOP/s (roofline denominator of the register-resident
 * trajectory kernels; the driver's MEASURED_PEAKS.json only holds HBM and bf16 tensor peaks) */
gmcmc_status gmcmc_measure_fp32_peak(gmcmc_ctx*, double* tflops);
/* queues about `approx_ms` (0 .. 100) milliseconds of all-SM FP32 FMA work on the context's stream and returns without
 * synchronising.  Benchmark plumbing: a B200 runs the first millisecond after an idle gap (a host synchronisation)
 * about 3 % slow, so a timed region that follows a barrier is preceded by this (bench.py; no reference counterpart). */
gmcmc_status gmcmc_ctx_warm_fp32(gmcmc_ctx*, double approx_ms);

/* ---- targets (distributions.rs traits Target / BatchedGradientTarget / GradientTarget :67-110) */
gmcmc_status gmcmc_target_create(gmcmc_ctx*, gmcmc_target_kind kind, gmcmc_dtype dtype, int dim,
                                 const double* params, size_t n_params, gmcmc_target** out);
/* Custom target = a user-written device log-density-and-gradient function compiled AHEAD OF TIME with nvcc
 * into a plugin shared library (general_mcmc_b200/csrc/gmcmc_custom_target.cuh, GMCMC_REGISTER_CUSTOM_TARGET);
 * ≙ a user impl of GradientTarget / BatchedGradientTarget (distributions.rs:67-90).  `params` are converted
 * to `dtype` and handed to the device function.  Usable with gmcmc_hmc_create and gmcmc_nuts_create. */
gmcmc_status gmcmc_target_create_custom(gmcmc_ctx*, const char* plugin_path, gmcmc_dtype dtype,
                                        const double* params, size_t n_params, gmcmc_target** out);
gmcmc_status gmcmc_target_destroy(gmcmc_target*);
/* logp and gradient for a batch of host positions [n, dim] (≙ BatchedHamiltonianTarget::logp_and_grad,
 * batched_hmc.rs:18-22).  grad_out may be NULL. */
gmcmc_status gmcmc_target_logp_grad(gmcmc_target*, const void* x_host, size_t n, void* logp_out,
                                    void* grad_out, gmcmc_math_mode mode);

/* ---- samplers ------------------------------------------------------------------------------ */
/* ≙ HMC::new (hmc.rs:113-134) / BatchedGenericHMC::new (batched_hmc.rs:62-84).
 * init_host: [n_chains, dim] row-major in the target's dtype.  chain_offset: global index of local
 * chain 0 (0 on a single GPU). */
gmcmc_status gmcmc_hmc_create(gmcmc_ctx*, gmcmc_target*, size_t n_chains, uint64_t chain_offset,
                              const void* init_host, double step_size, uint32_t n_leapfrog,
                              uint64_t seed, gmcmc_sampler** out);
/* ≙ MetropolisHastings::new + seed (metropolis_hastings.rs:151-197) with an IsotropicGaussian
 * proposal (distributions.rs:349-390). */
gmcmc_status gmcmc_mh_create(gmcmc_ctx*, gmcmc_target*, double proposal_std, size_t n_chains,
                             uint64_t chain_offset, const void* init_host, uint64_t seed,
                             gmcmc_sampler** out);
/* Integer-state Metropolis–Hastings: ≙ MetropolisHastings<S = i32, T = f64>::new + seed (metropolis_hastings.rs:151-197)
 * with the discrete targets and the +-1 random-walk proposals of the reference's discrete-state tests
 * (tests/metrohast_poisson_test.rs:18-86 Poisson(lambda), :195-252 Binomial(n, p)): unnorm_logp as written there (ln k!
 * summed term by term), proposal k +- 1 with probability 1/2 clamped to the support, Proposal::logp = ln 0.5 both ways.
 * State: int32 [n_chains, dim] (dim <= 8, coordinates independent); samples are f64 [C, n, dim] like every MH trace
 * (core.rs:34-51).  Runs through gmcmc_run / gmcmc_run_device / gmcmc_run_stats / gmcmc_step / gmcmc_positions
 * (int32) / gmcmc_counters_get like the other samplers.  RNG: stream 0 block 0, bit i of word 0 = direction of
 * coordinate i (1: +1); stream 1 block 0 = accept uniform. */
typedef enum { GMCMC_ITARGET_POISSON = 0 /* params [lambda] */, GMCMC_ITARGET_BINOMIAL = 1 /* params [n, p] */ } gmcmc_int_target_kind;
gmcmc_status gmcmc_mh_int_create(gmcmc_ctx*, gmcmc_int_target_kind kind, const double* params, size_t n_params,
                                 size_t n_chains, int dim, uint64_t chain_offset, const int32_t* init_host,
                                 uint64_t seed, gmcmc_sampler** out);
/* Test hook: the next n_steps transitions take these directions (int8 +1 / -1, [n_steps, n_chains, dim]) and
 * ln u values (f64 [n_steps, n_chains]); log ratio (f64) and decisions through gmcmc_read_diagnostics. */
gmcmc_status gmcmc_mh_int_inject(gmcmc_sampler*, const int8_t* steps, const double* ln_u, size_t n_steps);
/* Gibbs sampling: ≙ GibbsSampler::new + set_seed (gibbs.rs:139-162); one transition = one full sweep
 * (GibbsMarkovChain::step, gibbs.rs:89-105: for i in 0 .. dim, state[i] = conditional.sample(i, &state)).  The
 * reference's `Conditional<S>` is a host closure; here it is one of the conditionals of the reference's own tests
 * (gibbs.rs:177-245: ConstantConditional { c }, any dim <= 8; MixtureConditional { mu0, sigma0, mu1, sigma1, pi0 } on the
 * state [x, z]) or a device function compiled ahead of time into a plugin (general_mcmc_b200/csrc/
 * gmcmc_custom_conditional.cuh, GMCMC_REGISTER_CONDITIONAL; the plugin fixes dim).  State and samples: f64
 * [n_chains, dim] / [C, n, dim] (core.rs:34-51).  Runs through gmcmc_run / gmcmc_run_device / gmcmc_run_stats /
 * gmcmc_step / gmcmc_positions like the other samplers.  RNG: draw k (< 4) of coordinate i at transition s comes from
 * Philox block (chain, s, stream 0, block 4 i + k): words (0, 1) -> the normal, words (2, 3) -> the uniform.  Unlike the
 * reference, whose per-chain clones of a conditional share one RNG state (every chain then draws the same numbers,
 * gibbs.rs:145-148), chains here are independent. */
typedef enum { GMCMC_COND_CONSTANT = 0 /* params [c] */, GMCMC_COND_MIXTURE_XZ = 1 /* params [mu0, sigma0, mu1, sigma1, pi0] */ } gmcmc_conditional_kind;
gmcmc_status gmcmc_gibbs_create(gmcmc_ctx*, gmcmc_conditional_kind kind, const double* params, size_t n_params,
                                size_t n_chains, int dim, uint64_t chain_offset, const double* init_host,
                                uint64_t seed, gmcmc_sampler** out);
gmcmc_status gmcmc_gibbs_create_custom(gmcmc_ctx*, const char* plugin_path, const double* params, size_t n_params,
                                       size_t n_chains, uint64_t chain_offset, const double* init_host,
                                       uint64_t seed, gmcmc_sampler** out);
/* Test hook: the next n_steps sweeps take the first normal and the first uniform of every (sweep, coordinate) from
 * these f64 arrays [n_steps, n_chains, dim]. */
gmcmc_status gmcmc_gibbs_inject(gmcmc_sampler*, const double* normals, const double* uniforms, size_t n_steps);
/* ≙ NUTS::new (nuts.rs:156-190) / GenericNUTS::new (generic_nuts.rs:370-398), identity mass.
 * max_depth 0 = uncapped like the reference up to an internal safety cap of 20.
 * init_step_size <= 0: find_reasonable_epsilon (generic_nuts.rs:1025-1102) per chain. */
gmcmc_status gmcmc_nuts_create(gmcmc_ctx*, gmcmc_target*, size_t n_chains, uint64_t chain_offset,
                               const void* init_host, double target_accept, uint32_t max_depth,
                               double init_step_size, uint64_t seed, gmcmc_sampler** out);
gmcmc_status gmcmc_sampler_destroy(gmcmc_sampler*);

gmcmc_status gmcmc_set_seed(gmcmc_sampler*, uint64_t seed);  /* ≙ set_seed / seed */
gmcmc_status gmcmc_set_math_mode(gmcmc_sampler*, gmcmc_math_mode);
/* HMC step-size adaptation during the discard phase of the next run (dual averaging, constants of
 * generic_nuts.rs:638-641: gamma 0.05, t0 10, kappa 0.75, mu = ln(10 eps0)). */
gmcmc_status gmcmc_set_adaptation(gmcmc_sampler*, gmcmc_adapt_mode, double target_accept);
gmcmc_status gmcmc_set_step_size(gmcmc_sampler*, double step_size);

/* Test hook for per-step equivalence: the next `n_steps` transitions consume these host arrays
 * instead of Philox.  HMC: normals [n_steps, n_chains, dim], ln_u [n_steps, n_chains].
 * MH: same shapes (normals = proposal noise).  NUTS: see gmcmc_nuts_inject. */
gmcmc_status gmcmc_inject(gmcmc_sampler*, const void* normals, const void* ln_u, size_t n_steps);
/* Test hook for the production 2-D fast-mode MH kernel (which draws from Philox only): the next `n_steps`
 * transitions run through that same kernel and record, per step and chain, the log ratio and the decision (read
 * with gmcmc_read_diagnostics) and the proposal noise / accept uniform it actually used (gmcmc_mh_read_draws:
 * float [n_steps, n_chains, 3] = z0, z1, u).  Feeding those draws to the reference's MHMarkovChain::step
 * (metropolis_hastings.rs:306-318) must reproduce the chain. */
gmcmc_status gmcmc_mh_record(gmcmc_sampler*, size_t n_steps);
gmcmc_status gmcmc_mh_read_draws(gmcmc_sampler*, float* draws_out);
/* NUTS: per-chain streams in the reference's draw order: normals [C, n_norm],
 * exp1 [C, n_exp], unif [C, n_unif], all f64. */
gmcmc_status gmcmc_nuts_inject(gmcmc_sampler*, const double* normals, size_t n_norm, const double* exp1,
                               size_t n_exp, const double* unif, size_t n_unif);
/* Warm-up mass-matrix adaptation of NUTS.  ≙ GenericNUTS::new_with_mass_matrix(.., NUTSMassMatrixConfig)
 * (generic_nuts.rs:40-78, 379-398); the defaults of NUTSMassMatrixConfig are start_buffer 75, end_buffer 50,
 * initial_window 25, regularize 0.05, jitter 1e-6.  Call before the first run.  Each chain adapts its own
 * diagonal inverse mass from the running variance of its warm-up positions over doubling windows
 * (MassMatrixWarmup :134-175, RunningCov :81-132, maybe_update_mass_matrix :948-969); after every update a
 * fresh momentum (Philox stream 3 of the transition that ended the window) probes a new step size and dual
 * averaging restarts (:906-918).  GMCMC_MASS_DENSE ≙ MassMatrixAdaptation::Dense: the running covariance keeps the full
 * outer-product sums (RunningCov :81-132), a window end turns them into the regularised covariance, its Cholesky factor
 * and inverse per chain (dense_from_cov :208-226, cholesky_spd / invert_spd_from_cholesky :306-359, update :970-997), momenta
 * are chol z, velocities and kinetic energy use the dense inverse, and — as in the reference — the step-size search and the
 * sub-tree U-turn tests keep the identity.  Above dense_max_dim (default 75, gmcmc_nuts_set_dense_max_dim, call it first) the
 * reference keeps diagonal statistics but still asks for a dense update and therefore never updates (:612-617, :972-974):
 * the mass stays the identity, reproduced here.  Memory: 5 arrays of n_chains * dim^2 elements. */
typedef enum { GMCMC_MASS_NONE = 0, GMCMC_MASS_DIAGONAL = 1, GMCMC_MASS_DENSE = 2 } gmcmc_mass_adaptation;
gmcmc_status gmcmc_nuts_set_mass_adaptation(gmcmc_sampler*, gmcmc_mass_adaptation kind, size_t start_buffer,
                                            size_t end_buffer, size_t initial_window, double regularize,
                                            double jitter);
gmcmc_status gmcmc_nuts_set_dense_max_dim(gmcmc_sampler*, size_t dense_max_dim);   /* NUTSMassMatrixConfig::dense_max_dim, :50,76 */
/* Current inverse mass — diagonal adaptation: [C, dim]; dense adaptation: [C, dim, dim] (sampler dtype; identity before the
 * first update) — and the number of updates so far. */
gmcmc_status gmcmc_nuts_mass_matrix(gmcmc_sampler*, void* inv_mass_out, uint64_t* n_updates_out);
/* NUTS per-chain state (tests / diagnostics): current step sizes [C] (sampler dtype), accumulated
 * leapfrogs [C], consumed injected draws [C][3] (normals, exp1, unif).  Any pointer may be NULL. */
gmcmc_status gmcmc_nuts_state(gmcmc_sampler*, void* eps_out, long long* leapfrogs_out,
                              unsigned long long* used_out);
/* Per-step diagnostics of the injected transitions (HMC/MH): log_accept [n_steps, C] (sampler dtype),
 * accepted [n_steps, C] (u8), prop_q / prop_p [n_steps, C, dim] (HMC only; end of trajectory).
 * Any pointer may be NULL.  Valid after the run that consumed the injection. */
gmcmc_status gmcmc_read_diagnostics(gmcmc_sampler*, void* log_accept, uint8_t* accepted, void* prop_q,
                                    void* prop_p);

/* ≙ HMC::step (hmc.rs:308), MarkovChain::step (core.rs:79-85): one transition, nothing recorded.  For an HMC sampler with
 * step-size adaptation enabled (gmcmc_set_adaptation, an extension the reference's HMC does not have) the step counts as a
 * warm-up transition and advances dual averaging; call gmcmc_set_adaptation(s, GMCMC_ADAPT_NONE, ...) first for a plain step. */
gmcmc_status gmcmc_step(gmcmc_sampler*);
/* ≙ HMC::run (hmc.rs:164-181), BatchedGenericHMC::run (batched_hmc.rs:93-112), ChainRunner::run
 * (core.rs:219-229), NUTS::run (nuts.rs:214-257).  out_host: [C, n_collect, dim] of out_dtype. */
gmcmc_status gmcmc_run(gmcmc_sampler*, size_t n_collect, size_t n_discard, void* out_host,
                       gmcmc_dtype out_dtype);
/* ≙ BatchedGenericHMC::run_positions (batched_hmc.rs:115-123): samples stay on the device.
 * *out_dev: library-owned [C, n_collect, dim] (sampler dtype; f64 for MH). */
gmcmc_status gmcmc_run_device(gmcmc_sampler*, size_t n_collect, size_t n_discard, void** out_dev);
/* Sizes the library-owned device sample buffer for runs of up to n_collect draws per chain now (a multi-GB
 * cudaMalloc) instead of inside the first gmcmc_run_device / gmcmc_run_stats that needs it.  n_collect = 0 frees it. */
gmcmc_status gmcmc_reserve_samples(gmcmc_sampler*, size_t n_collect);
/* ≙ run_progress (hmc.rs:245-306, core.rs:251-403, generic_nuts.rs:414-548) without the terminal UI:
 * samples (optional, may be NULL) + RunStats computed on the device over ALL ranks' chains.
 * Limits of the device statistics (checked before any transition is taken, so a refused call leaves the sampler
 * untouched): 4 <= n_collect <= 16385 (padded FFT length <= 16384) — GMCMC_ERR_INVALID / GMCMC_ERR_UNSUPPORTED
 * otherwise.  The same limits apply to gmcmc_split_rhat_ess / gmcmc_run_stats_from. */
gmcmc_status gmcmc_run_stats(gmcmc_sampler*, size_t n_collect, size_t n_discard, void* out_host_or_null,
                             gmcmc_dtype out_dtype, gmcmc_run_stats_t* stats);
gmcmc_status gmcmc_positions(gmcmc_sampler*, void* out_host); /* ≙ positions() hmc.rs:318 */
/* overwrite the chains' current positions from a host [C, dim] array (restart from new initial
 * positions without rebuilding the sampler; ≙ constructing HMC::new with other initial_positions) */
gmcmc_status gmcmc_set_positions(gmcmc_sampler*, const void* init_host);
gmcmc_status gmcmc_counters_get(gmcmc_sampler*, gmcmc_counters* out);
gmcmc_status gmcmc_sampler_info(gmcmc_sampler*, size_t* n_chains, int* dim, gmcmc_dtype* dtype);

/* ---- diagnostics (stats.rs:439-450 split_rhat_mean_ess) -------------------------------------- */
/* samples: [C, n, p] of `dtype`, host (on_device = 0) or device pointer.  rhat/ess: host float[p].
 * rhat is the reference orientation sqrt(W/var_hat).  With a distributed context, C is this rank's
 * shard and the result covers all ranks' chains. */
gmcmc_status gmcmc_split_rhat_ess(gmcmc_ctx*, const void* samples, size_t C, size_t n, size_t p,
                                  gmcmc_dtype dtype, int on_device, float* rhat, float* ess);
/* ≙ RunStats::from (stats.rs:383-394) */
gmcmc_status gmcmc_run_stats_from(gmcmc_ctx*, const void* samples, size_t C, size_t n, size_t p,
                                  gmcmc_dtype dtype, int on_device, gmcmc_run_stats_t* out);

/* Progress tracker.  ≙ MultiChainTracker::{new, step (once per draw), rhat, max_rhat, p_accept} (stats.rs:199-339):
 * what run_progress shows while sampling — per-chain running mean / mean of squares (f32, the reference's
 * recurrences), the EMA acceptance rate (alpha = 0.01, "row changed" test, folded over the chains of a step, step
 * after step) and the tracker R-hat sqrt(var_hat / W) — evaluated on the device from the draws [C, n, p] collected
 * so far.  rhat: host float[p]; max_rhat, p_accept: host scalars; any may be NULL.  Needs C >= 2, n >= 2.  With a
 * distributed context the figures cover THIS rank's chains (a progress display, not a diagnostic: use
 * gmcmc_split_rhat_ess / gmcmc_run_stats_from for the all-ranks R-hat and ESS). */
gmcmc_status gmcmc_tracker_stats(gmcmc_ctx*, const void* samples, size_t C, size_t n, size_t p,
                                 gmcmc_dtype dtype, int on_device, float* rhat, float* max_rhat, float* p_accept);

/* Column export of a sample tensor, ≙ the column builders of io::csv::save_csv / save_csv_tensor (io/csv.rs:47-147),
 * io::arrow::save_arrow (io/arrow.rs:53-117), io::parquet::save_parquet / save_parquet_tensor (io/parquet.rs:49-222):
 * one row per (chain, observation) pair with columns chain:u32, observation:u32, dim_0 .. dim_{d-1}:f64.
 * GMCMC_ROWS_CHAIN_MAJOR: chain outer, observation inner (save_csv, save_arrow, save_parquet);
 * GMCMC_ROWS_OBS_MAJOR:   observation outer, chain inner (save_parquet_tensor's [obs, chain, dim] order, parquet.rs:166-176).
 * samples: [C, n, d] of `dtype`, device (on_device = 1, e.g. the pointer gmcmc_run_device returned) or host.  The
 * transpose and the widening to f64 run on the device.  chain_out / obs_out: host u32 [C n] (chain values start at
 * chain_base: the rank's chain offset); dims_out: host f64 [d][C n] (column k contiguous).  Any output may be NULL. */
typedef enum { GMCMC_ROWS_CHAIN_MAJOR = 0, GMCMC_ROWS_OBS_MAJOR = 1 } gmcmc_row_order;
gmcmc_status gmcmc_export_columns(gmcmc_ctx*, const void* samples, size_t C, size_t n, size_t d, gmcmc_dtype dtype,
                                  int on_device, gmcmc_row_order order, uint32_t chain_base, uint32_t* chain_out,
                                  uint32_t* obs_out, double* dims_out);

/* raw Philox4x32-10 blocks computed on the device (contract check): ctr [n,4], key [2] -> out [n,4] */
gmcmc_status gmcmc_philox_blocks(gmcmc_ctx*, const uint32_t* ctr_host, size_t n, const uint32_t* key,
                                 uint32_t* out_host);

const char* gmcmc_last_error(void);
const char* gmcmc_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GMCMC_H */
