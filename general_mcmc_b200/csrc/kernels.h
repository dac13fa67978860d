// kernels.h — type-erased launch descriptors shared by the host runtime (runtime.cu) and the
// kernel translation units (hmc_*.cu, mh.cu, stats.cu, nuts.cu).  Internal; the public surface
// is include/gmcmc.h.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace gm {

constexpr int kMaxScalarParams = 8;
#ifndef GM_HMC_BLOCK
// tuning knob.  Measured on B200: 64-thread CTAs run K1 exactly as fast as 128-thread ones (profiles/r2_k1_cta_size_experiment.txt),
// and a 13th resident warp cannot be had from smaller CTAs: registers are granted per SM sub-partition, so 13 or 18 warps per SM
// get the register budget of 16 or 20 (ptxas: 128 / 96 registers with spills where 12 / 16 warps get 168 / 128)
#define GM_HMC_BLOCK 128
#endif
constexpr int kHmcBlock = GM_HMC_BLOCK;  // threads per CTA of the trajectory kernels

struct TargetDesc {
  int kind;                          // gmcmc_target_kind
  int dtype;                         // gmcmc_dtype
  int dim;
  double sp[kMaxScalarParams];       // small parameter block (cast to T in the kernel)
  const void* dparams;               // device parameter block in T (dense Gaussian, mixture), or null
  int n_comp;                        // mixture components
};

struct HmcLaunch {
  TargetDesc tgt;
  size_t n_chains;
  uint64_t chain_offset;
  uint64_t seed;
  uint32_t step_base;    // transition index of the first transition of this launch
  void* positions;       // [C, d] T, in/out
  const void* eps;       // device step size(s), T; null: eps_val is used
  double eps_val;        // step size passed by value (fixed step size known on the host)
  int eps_stride;        // 0: one shared scalar, 1: per chain
  uint32_t n_leapfrog;
  uint32_t n_steps;      // transitions in this launch
  uint32_t n_skip;       // the first n_skip transitions are not recorded
  void* out;             // samples base [C, out_n, d] T (may be null when n_steps == n_skip)
  size_t out_n;          // samples per chain in `out`
  uint32_t out_t0;       // sample slot of the first recorded transition of this launch
  // statistics
  unsigned long long* accept_total;  // [1]
  unsigned long long* diverge_total; // [1]
  double* alpha_part;                // [n_steps][warps of the grid] per-transition, per-warp sums of min(1, exp(log_accept)), or null
  // per-chain dual averaging (GMCMC_ADAPT_PER_CHAIN), all T [C]; null otherwise
  void* da_eps; void* da_eps_bar; void* da_h_bar; void* da_mu;
  uint32_t da_m_base;    // adaptation iteration of the first transition of this launch (1-based m = base + s + 1)
  uint32_t da_n_adapt;   // adapt while m <= da_n_adapt
  double da_delta;
  // injection (parity mode) / diagnostics; null when unused
  const void* inj_normals;  // [n, C, d]
  const void* inj_lnu;      // [n, C]
  void* diag_logacc;        // [n, C]
  uint8_t* diag_acc;        // [n, C]
  void* diag_pq;            // [n, C, d]
  void* diag_pp;            // [n, C, d]
  // decomposition: a chain is spread over `lpc` lanes (power of two <= 32), `epl` elements per lane
  int epl, lpc;
};

struct EvalLaunch {  // logp + grad of a batch of points (gmcmc_target_logp_grad)
  TargetDesc tgt;
  size_t n;
  const void* x;   // [n, d]
  void* logp;      // [n]
  void* grad;      // [n, d] or null
  int epl, lpc;
};

struct MhLaunch {
  TargetDesc tgt;
  double prop_std;
  size_t n_chains;
  uint64_t chain_offset;
  uint64_t seed;
  uint32_t step_base;
  void* state;        // [C, d] T in/out
  uint32_t n_steps, n_skip;
  double* out;        // [C, out_n, d] f64
  size_t out_n;
  uint32_t out_t0;
  unsigned long long* accept_total;
  const void* inj_normals;  // [n, C, d] T
  const void* inj_lnu;      // [n, C] T
  void* diag_logratio;      // [n, C] T
  uint8_t* diag_acc;        // [n, C]
  float* diag_draws;        // [n, C, 3] draws used by the 2-D fast kernel (gmcmc_mh_record), or null
};

// integer-state MH (mh_int.cu): Poisson / Binomial targets, +-1 random-walk proposal (tests/metrohast_poisson_test.rs)
struct MhIntLaunch {
  int kind, dim, n;                 // 0 Poisson, 1 Binomial(n, p)
  double lambda, ln_lambda, ln_p, ln_1mp, ln_half;   // logarithms evaluated on the host (glibc), like the CPU reference
  const double* lnfact; int n_tab;  // device table of ln(k!), k < n_tab
  size_t n_chains;
  uint64_t chain_offset, seed;
  uint32_t step_base, n_steps, n_skip;
  int* state;                       // [C, d] int32 in/out
  double* out; size_t out_n; uint32_t out_t0;
  unsigned long long* accept_total;
  const signed char* inj_steps;     // [n, C, d] +1 / -1, or null
  const double* inj_lnu;            // [n, C], or null
  double* diag_logratio; uint8_t* diag_acc;
};
cudaError_t launch_mh_int(const MhIntLaunch&, cudaStream_t);
int mh_int_max_dim();

// Gibbs sweeps (gibbs_kernel.cuh / gibbs.cu): built-in conditionals of the reference's tests, or a plugin's
struct GibbsLaunch {
  int kind, dim;                    // 0 ConstantConditional [c], 1 MixtureConditional [mu0, sigma0, mu1, sigma1, pi0] on [x, z]
  const double* params;             // device
  size_t n_chains;
  uint64_t chain_offset, seed;
  uint32_t step_base, n_steps, n_skip;
  double* state;                    // [C, d] in/out
  double* out; size_t out_n; uint32_t out_t0;
  const double* inj_normals;        // [n, C, d] first normal of each (transition, coordinate), or null
  const double* inj_uniforms;       // [n, C, d] first uniform, or null
};
cudaError_t launch_gibbs(const GibbsLaunch&, cudaStream_t);
int gibbs_max_dim();
struct CustomConditionalVTable {
  int abi_version;
  int dim;
  cudaError_t (*launch_gibbs)(const GibbsLaunch&, cudaStream_t);
};

// K3: dense-Gaussian HMC with the gradient GEMM on tcgen05 (dense_tc.cu)
struct DenseTc;
struct DenseTcStep {
  void* q;                 // [C, d] f32 current positions, in/out
  uint64_t chain_offset, seed;
  uint32_t step;           // Philox transition index
  double eps;
  uint32_t n_leapfrog;
  void* out; size_t out_n; long long slot;     // slot < 0: transition not recorded
  unsigned long long* accept_total; unsigned long long* diverge_total;
  const void* inj_normals; const void* inj_lnu;          // this transition's slices, or null
  void* diag_logacc; uint8_t* diag_acc; void* diag_pq; void* diag_pp;
};
DenseTc* dense_tc_create(size_t n_chains, int d, const double* params, const char** err);
void dense_tc_destroy(DenseTc*);
int dense_tc_transition(DenseTc*, const DenseTcStep&, cudaStream_t);

struct NutsLaunch {
  TargetDesc tgt;
  int init_only;          // 1: run init_chain_state (momentum draw + find_reasonable_epsilon + mu) instead of transitions
  size_t n_chains;
  uint64_t chain_offset;
  uint64_t seed;
  uint32_t step_base;
  void* positions;
  uint32_t n_steps, m_base, n_discard;
  long long rec_off;
  int write_init;
  void* out;
  size_t out_n;
  void* eps; void* eps_bar; void* h_bar; void* mu;   // [C] T
  double target_accept;
  int max_depth;          // effective cap
  void* ws_edges; void* ws_first; void* ws_prime;
  int cap;
  unsigned long long* leapfrog_total; unsigned long long* diverge_total; unsigned long long* depth_total;
  unsigned long long* accept_total;
  long long* chain_leapfrogs;
  const double* inj_normals; size_t n_norm;
  const double* inj_exp1; size_t n_exp;
  const double* inj_unif; size_t n_unif;
  unsigned long long* inj_used;
  unsigned long long* queue;    // [1] device counter of the dynamic chain queue
  int epl, lpc;
  // diagonal mass matrix (GenericNUTS::new_with_mass_matrix, generic_nuts.rs:379-398); null = identity
  void* mass_inv; void* mass_sqrt;                       // [C, d] T: 1 / var, sqrt(var)
  void* run_mean; void* run_m2;                          // RunningCov of the warm-up window, [C, d], [C, d]
  uint32_t run_n_base;                                   // its sample count when the launch starts (same for every chain)
  uint32_t collect_after, collect_before;                // positions are collected for collect_after < m < collect_before
  int probe;              // with init_only: mass-matrix update probe (generic_nuts.rs:906-918) instead of init_chain_state
  // dense mass matrix (MassMatrix::Dense): inverse mass [C, d, d], lower Cholesky factor [C, d, d], running outer-product
  // sums [C, d, d]; dense_active = 0 until the first update (identity arithmetic)
  void* mass_dinv; void* mass_chol; void* run_m2d; int dense_active;
};

// mass_dense.cu: maybe_update_mass_matrix, Dense branch (generic_nuts.rs:970-997) + dense_from_cov (:208-226) for every chain
struct DenseMassUpdate {
  int dtype; size_t n_chains; int d; unsigned int n;     // n = positions in the running covariance
  void* run_mean; void* run_m2; void* run_m2d;           // reset on return
  void* inv; void* chol;                                 // [C, d, d] out
  void* scratch_l; void* scratch_invl;                   // [d, d, C] (chain fastest) work arrays
  double regularize, jitter;
  int* state;                                            // [C]: 0 identity, 1 all-ones diagonal (fallback), 2 dense
};
cudaError_t launch_dense_mass_update(const DenseMassUpdate&, cudaStream_t);

struct StatsLaunch {
  const void* samples;   // [C, n, p] device, f32 (dtype 0) or f64 (dtype 1)
  int dtype;
  size_t C, n;
  int p;
  size_t N;              // padded FFT length (power of two >= 2*(n/2)-1)
  int log2n;
  int ppb;               // parameters per CTA of stats_accumulate
  int n_groups;          // chain groups (grid.x of stats_accumulate)
  const void* tw;        // [N/2] float2 twiddles
  float* part_spec;      // [n_groups, p, N/2+1]
  double* part_mom;      // [n_groups, p, 3]
  float* spec;           // [p, N/2+1]   chain-summed power spectrum (all-reduced across ranks)
  double* mom;           // [p, 3]       sum of split-chain means, of their squares, of within variances
  float* rhat;           // [p] reference orientation sqrt(W / var_hat)
  float* rhat_std;       // [p] sqrt(var_hat / W)
  float* ess;            // [p]
  float* acov;           // [p, n/2] mean autocovariance (optional, may be null)
};

// custom-target plugins (gmcmc_custom_target.cuh): the launchers a plugin instantiates for its target
constexpr int kCustomAbiVersion = 3;
constexpr int kTargetCustom = 7;      // TargetDesc.kind of a plugin target
struct CustomTargetVTable {
  int abi_version;
  int dim;
  cudaError_t (*launch_hmc)(const HmcLaunch&, cudaStream_t);
  cudaError_t (*launch_eval)(const EvalLaunch&, cudaStream_t);
  cudaError_t (*launch_nuts)(const NutsLaunch&, cudaStream_t);
  cudaError_t (*launch_mh)(const MhLaunch&, cudaStream_t);
};

// export.cu: columns [kbase, kbase + d) of [C, n, ld] samples -> chain:u32 [rows], observation:u32 [rows], dims:f64 [d][rows]
cudaError_t launch_export_columns(const void* samples, int dtype, size_t C, size_t n, int ld, int kbase, int d, int obs_major,
                                  unsigned int chain_base, unsigned int* chain_col, unsigned int* obs_col, double* dims, cudaStream_t st);

size_t stats_npad(size_t n);
int stats_ppb(size_t N);
cudaError_t launch_tracker(const void* samples, int dtype, size_t C, size_t n, int p, float* mean /*[C,p]*/, float* mean_sq /*[C,p]*/,
                           float* rhat /*[p]*/, float* p_accept /*[1]*/, cudaStream_t st);   // K6, stats.cu
int stats_groups(size_t N, size_t p, size_t C, int sm_count);   // chain groups of the accumulate kernel
void stats_fill_twiddles(size_t N, float* host_tw);
cudaError_t launch_stats_accumulate(const StatsLaunch&, cudaStream_t);
cudaError_t launch_stats_reduce(const StatsLaunch&, cudaStream_t);
cudaError_t launch_stats_finalize(const StatsLaunch&, double total_chains, const double* total_chains_dev, cudaStream_t);

// launchers (each returns the cudaError of the launch).  *_exact are compiled with --fmad=false.
cudaError_t launch_hmc_fast(const HmcLaunch&, cudaStream_t);
cudaError_t launch_hmc_exact(const HmcLaunch&, cudaStream_t);
cudaError_t launch_eval_fast(const EvalLaunch&, cudaStream_t);
cudaError_t launch_eval_exact(const EvalLaunch&, cudaStream_t);
cudaError_t launch_mh_fast(const MhLaunch&, cudaStream_t);
cudaError_t launch_mh_exact(const MhLaunch&, cudaStream_t);

cudaError_t launch_nuts_fast(const NutsLaunch&, cudaStream_t);
cudaError_t launch_nuts_exact(const NutsLaunch&, cudaStream_t);
bool choose_nuts_decomposition(int dim, int dtype, int kind, int* epl, int* lpc);
constexpr int kNutsDepthCapHost = 20;

// picks (epl, lpc) for a dimension; returns false if unsupported by the register-resident kernels
bool choose_decomposition(int dim, int dtype, int kind, int* epl, int* lpc);

}  // namespace gm
