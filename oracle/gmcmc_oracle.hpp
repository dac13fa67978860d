// gmcmc_oracle.hpp — CPU restatement of the general-mcmc many-chain sampling hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
// link, import or execute it, and only as the checker / reported CPU baseline.
//
// Provenance: the reference (SauersML/general-mcmc 0.8.0, /root/reference) is pure Rust and
// cannot be compiled in this image (no cargo/rustc, no Cargo.lock, no vendored crates).  Every
// function below restates one reference function and cites its file:line.  Arithmetic follows
// the reference's operation order with sequential left-to-right sums; compile with
// -ffp-contract=off (Rust never contracts `a + b*alpha`, euclidean.rs:93-97).
//
// Pinning: the RNG-free known-answer tests of the reference (nuts.rs:508-601, stats.rs:734-839,
// distributions.rs:580-614, 820-839, generic_nuts.rs:1427-1455) are reproduced in
// tests/test_oracle_kat.py against this file.  What no reference test pins (a full HMC/MH step,
// split_rhat_mean_ess values, any seeded sample, the bit order of ndarray/burn reductions and
// of burn's autodiff gradient) is "parity unpinned" and rests on review against the cited lines.
//
// Randomness is injected: every sampler takes arrays of N(0,1) draws, ln(u) values, Exp(1)
// draws and uniforms, consumed in the reference's draw order, so that the CUDA path and this
// file see identical numbers (the reference's Xoshiro256++/ziggurat streams are third-party
// code absent from /root/reference: rand 0.9, rand_distr 0.5 — RNG-stream parity unpinned).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

namespace orc {

// Mirrors include/gmcmc.h gmcmc_target_kind (duplicated so the oracle stays standalone).
enum TargetKind : int {
  T_ISO_GAUSS = 0,      // distributions.rs:398-406 (+ gradient -x/std^2; nuts.rs:484-496 StandardNormal)
  T_GAUSS2D = 1,        // distributions.rs:191-208 (MH target, no gradient in the reference)
  T_DIFF_GAUSS2D = 2,   // distributions.rs:215-291
  T_DENSE_GAUSS = 3,    // N-D generalisation of DiffableGaussian2D
  T_ROSENBROCK2D = 4,   // distributions.rs:495-515
  T_ROSENBROCK_ND = 5,  // distributions.rs:535-555
  T_GAUSS_MIXTURE = 6   // synthetic
};

// ------------------------------------------------------------------------------------------
// Targets.  All take params as doubles (cast to T once, like the reference's constructors).
// ------------------------------------------------------------------------------------------
template <class T>
struct Target {
  int kind = T_ROSENBROCK_ND;
  int dim = 0;
  std::vector<T> p;  // parameter block in T

  // derived (DiffableGaussian2D::new, distributions.rs:229-253)
  T inv_cov[2][2] = {{0, 0}, {0, 0}};
  T norm_const = 0;

  Target() {}
  Target(int kind_, int dim_, const double* params, size_t n) : kind(kind_), dim(dim_) {
    p.resize(n);
    for (size_t i = 0; i < n; ++i) p[i] = (T)params[i];
    if (kind == T_DIFF_GAUSS2D) {
      // distributions.rs:229-253
      T c00 = p[2], c01 = p[3], c10 = p[4], c11 = p[5];
      T det_cov = c00 * c11 - c01 * c10;
      T inv_det = T(1) / det_cov;
      inv_cov[0][0] = c11 * inv_det;
      inv_cov[0][1] = -c01 * inv_det;
      inv_cov[1][0] = -c10 * inv_det;
      inv_cov[1][1] = c00 * inv_det;
      T logdet = std::log(det_cov);
      T two = T(1) + T(1);
      const T pi = (T)3.14159265358979323846264338327950288;
      norm_const = -(two * std::log(two * pi) + logdet) / two;
    }
  }

  // unnormalised log density only (MH path: Target::unnorm_logp, distributions.rs:107-110)
  T logp(const T* x) const {
    switch (kind) {
      case T_ISO_GAUSS: {  // distributions.rs:398-406
        T sum = 0;
        for (int i = 0; i < dim; ++i) sum = sum + x[i] * x[i];
        return -T(0.5) * sum / (p[0] * p[0]);
      }
      case T_GAUSS2D: {  // distributions.rs:195-207
        T a = p[2], b = p[3], c = p[4], d = p[5];
        T det = a * d - b * c;
        T d0 = x[0] - p[0], d1 = x[1] - p[1];
        T i00 = d / det, i01 = (-b) / det, i10 = (-c) / det, i11 = a / det;
        // diff.dot(&inv_cov) : row vector, then .dot(&diff)
        T v0 = d0 * i00 + d1 * i10;
        T v1 = d0 * i01 + d1 * i11;
        return -T(0.5) * (v0 * d0 + v1 * d1);
      }
      default: {
        std::vector<T> g(dim);
        return logp_and_grad(x, g.data());
      }
    }
  }

  // HamiltonianTarget::logp_and_grad (generic_hmc.rs:14-17).  The reference obtains the gradient
  // by burn reverse-mode autodiff (hmc.rs:42-61, distributions.rs:83-89); here it is the analytic
  // gradient written in the order the backward pass accumulates it (parity unpinned at the bit
  // level; pinned to 1e-5 by nuts.rs:521-586 for DiffableGaussian2D).
  T logp_and_grad(const T* x, T* g) const {
    const int d = dim;
    switch (kind) {
      case T_ISO_GAUSS: {
        T var = p[0] * p[0];
        T sum = 0;
        for (int i = 0; i < d; ++i) {
          sum = sum + x[i] * x[i];
          g[i] = -x[i] / var;
        }
        return -T(0.5) * sum / var;
      }
      case T_GAUSS2D: {
        // gradient of distributions.rs:195-207 (not used by the reference; for completeness)
        T a = p[2], b = p[3], c = p[4], dd = p[5];
        T det = a * dd - b * c;
        T d0 = x[0] - p[0], d1 = x[1] - p[1];
        T i00 = dd / det, i01 = (-b) / det, i10 = (-c) / det, i11 = a / det;
        T v0 = d0 * i00 + d1 * i10;
        T v1 = d0 * i01 + d1 * i11;
        T w0 = i00 * d0 + i01 * d1;
        T w1 = i10 * d0 + i11 * d1;
        g[0] = -T(0.5) * (v0 + w0);
        g[1] = -T(0.5) * (v1 + w1);
        return -T(0.5) * (v0 * d0 + v1 * d1);
      }
      case T_DIFF_GAUSS2D: {
        // distributions.rs:265-291: delta = x - mean; z = delta . inv_cov; quad = sum(z*delta);
        // logp = norm_const - 0.5*quad.   backward: g = (-0.5*z) + (-0.5*delta) . inv_cov^T
        T d0 = x[0] - p[0], d1 = x[1] - p[1];
        T z0 = d0 * inv_cov[0][0] + d1 * inv_cov[1][0];
        T z1 = d0 * inv_cov[0][1] + d1 * inv_cov[1][1];
        T quad = z0 * d0 + z1 * d1;
        T h = T(0.5);
        T m0 = -(h * d0), m1 = -(h * d1);
        g[0] = (-(h * z0)) + (m0 * inv_cov[0][0] + m1 * inv_cov[0][1]);
        g[1] = (-(h * z1)) + (m0 * inv_cov[1][0] + m1 * inv_cov[1][1]);
        return norm_const - quad * h;
      }
      case T_DENSE_GAUSS: {
        // params: mu[d], P[d*d] row-major (symmetric), norm_const.  z = delta . P ; grad = -z
        const T* mu = p.data();
        const T* P = p.data() + d;
        T nc = p[(size_t)d + (size_t)d * d];
        std::vector<T> delta(d);
        for (int i = 0; i < d; ++i) delta[i] = x[i] - mu[i];
        T quad = 0;
        for (int j = 0; j < d; ++j) {
          T z = 0;
          for (int i = 0; i < d; ++i) z = z + delta[i] * P[(size_t)i * d + j];
          g[j] = -z;
          quad = quad + z * delta[j];
        }
        return nc - quad * T(0.5);
      }
      case T_ROSENBROCK2D: {
        // distributions.rs:502-515: -( (a-x)^2 + b*(y-x^2)^2 )
        T a = p[0], b = p[1];
        T u = (-x[0]) + a;
        T t = x[1] - x[0] * x[0];
        T term1 = u * u;
        T term2 = (t * t) * b;
        // backward: d term2/dt = b*2*t ; dt/dx = -2x ; d term1/dx = -2u
        T gt = -((T(2) * t) * b);          // d(-term2)/dt
        g[1] = gt;
        g[0] = (T(2) * u) + (-(T(2) * (gt * x[0])));
        return -(term1 + term2);
      }
      case T_ROSENBROCK_ND: {
        // distributions.rs:544-554 (same body: examples/rosenbrock3d_hmc.rs:28-42)
        //   low = x[0..n-1], high = x[1..n]
        //   term_1 = (high - low^2)^2 * 100 ; term_2 = (1 - low)^2 ; logp = -sum(term_1 + term_2)
        T s = 0;
        T gt_prev = 0;
        for (int i = 0; i < d; ++i) {
          T gi = 0;
          if (i < d - 1) {
            T t = x[i + 1] - x[i] * x[i];
            T u = (-x[i]) + T(1);
            s = s + ((t * t) * T(100) + u * u);
            T gt = T(-200) * t;  // d(-term_1)/dt
            gi = (T(-2) * (gt * x[i])) + T(2) * u;
            if (i > 0) gi = gi + gt_prev;
            gt_prev = gt;
          } else if (i > 0) {
            gi = gt_prev;
          }
          g[i] = gi;
        }
        return -s;
      }
      case T_GAUSS_MIXTURE: {
        // params: K, sigma, w[K], mu[K*d].  logp = logsumexp_k( ln w_k - |x-mu_k|^2/(2 sigma^2) )
        int K = (int)p[0];
        T sigma = p[1];
        T inv_var = T(1) / (sigma * sigma);
        const T* w = p.data() + 2;
        const T* mu = p.data() + 2 + K;
        std::vector<T> a(K);
        T amax = -std::numeric_limits<T>::infinity();
        for (int k = 0; k < K; ++k) {
          T sq = 0;
          for (int i = 0; i < d; ++i) {
            T df = x[i] - mu[(size_t)k * d + i];
            sq = sq + df * df;
          }
          a[k] = std::log(w[k]) - T(0.5) * sq * inv_var;
          amax = std::max(amax, a[k]);
        }
        T se = 0;
        for (int k = 0; k < K; ++k) {
          a[k] = std::exp(a[k] - amax);
          se = se + a[k];
        }
        for (int i = 0; i < d; ++i) {
          T acc = 0;
          for (int k = 0; k < K; ++k) acc = acc + (a[k] / se) * (mu[(size_t)k * d + i] - x[i]);
          g[i] = acc * inv_var;
        }
        return amax + std::log(se);
      }
    }
    return 0;
  }
};

// ------------------------------------------------------------------------------------------
// Vector helpers (euclidean.rs:58-138)
// ------------------------------------------------------------------------------------------
template <class T>
inline void add_scaled_assign(T* a, const T* b, T alpha, int d) {  // euclidean.rs:93-97
  for (int i = 0; i < d; ++i) a[i] = a[i] + b[i] * alpha;
}
template <class T>
inline T dot(const T* a, const T* b, int d) {  // euclidean.rs:103-105 (order: sequential)
  T s = 0;
  for (int i = 0; i < d; ++i) s = s + a[i] * b[i];
  return s;
}

// ------------------------------------------------------------------------------------------
// HMC.  One transition of one chain.
//   generic_hmc.rs:166-221 (GenericHMC::step + leapfrog_chain, L+1 gradient evaluations) and
//   batched_hmc.rs:129-190 (BatchedGenericHMC::step + leapfrog, L+2 evaluations, two of them at
//   the same point): per chain both compute the same numbers, so one restatement serves both;
//   accept rule `ln_u <= log_accept` (generic_hmc.rs:198) == `log_accept >= ln_u`
//   (euclidean.rs:258-260, 527-533).
// ------------------------------------------------------------------------------------------
template <class T>
struct HmcStepInfo {
  T logp_current, logp_proposed, ke_current, ke_proposed, log_accept;
  int accepted;
};

template <class T>
HmcStepInfo<T> hmc_step(const Target<T>& tgt, T* q, const T* momentum, T ln_u, T step_size,
                        int n_leapfrog, T* prop_q_out /*nullable*/, T* prop_p_out /*nullable*/) {
  const int d = tgt.dim;
  std::vector<T> grad(d, T(0)), pq(q, q + d), pp(momentum, momentum + d);
  HmcStepInfo<T> r;
  const T ke_half = T(0.5);
  r.logp_current = tgt.logp_and_grad(q, grad.data());                 // generic_hmc.rs:174
  r.ke_current = dot(momentum, momentum, d) * ke_half;                // :178
  // leapfrog_chain, generic_hmc.rs:204-221
  const T half = T(0.5) * step_size;                                  // :213
  T logp = r.logp_current;
  for (int l = 0; l < n_leapfrog; ++l) {
    add_scaled_assign(pp.data(), grad.data(), half, d);               // :215
    add_scaled_assign(pq.data(), pp.data(), step_size, d);            // :216
    logp = tgt.logp_and_grad(pq.data(), grad.data());                 // :217
    add_scaled_assign(pp.data(), grad.data(), half, d);               // :218
  }
  r.logp_proposed = logp;
  r.ke_proposed = dot(pp.data(), pp.data(), d) * ke_half;             // :195
  r.log_accept = (r.logp_proposed - r.logp_current) + (r.ke_current - r.ke_proposed);  // :196
  r.accepted = (ln_u <= r.log_accept) ? 1 : 0;                        // :198
  if (prop_q_out) std::copy(pq.begin(), pq.end(), prop_q_out);
  if (prop_p_out) std::copy(pp.begin(), pp.end(), prop_p_out);
  if (r.accepted) std::copy(pq.begin(), pq.end(), q);                 // :199
  return r;
}

// ------------------------------------------------------------------------------------------
// Metropolis–Hastings with IsotropicGaussian proposal.
//   metropolis_hastings.rs:306-318 (MHMarkovChain::step), distributions.rs:368-390
// ------------------------------------------------------------------------------------------
template <class T>
inline T iso_proposal_logp(const T* from, const T* to, int d, T std_) {  // distributions.rs:378-390
  T lp = 0;
  T dd = (T)d;
  T two = T(2);
  T var = std_ * std_;
  for (int i = 0; i < d; ++i) {
    T diff = to[i] - from[i];
    T exponent = -(diff * diff) / (two * var);
    lp += exponent;
  }
  const T pi = (T)3.14159265358979323846264338327950288;
  lp += -dd * T(0.5) * std::log(var * pi * std_ * std_);
  return lp;
}

template <class T>
struct MhStepInfo {
  T current_lp, proposed_lp, log_accept_ratio;
  int accepted;
};

template <class T>
MhStepInfo<T> mh_step(const Target<T>& tgt, T prop_std, T* x, const T* normals, T ln_u) {
  const int d = tgt.dim;
  std::vector<T> proposed(d);
  for (int i = 0; i < d; ++i) proposed[i] = x[i] + normals[i] * prop_std;  // distributions.rs:368-376
  MhStepInfo<T> r;
  r.current_lp = tgt.logp(x);                                              // mh.rs:308
  r.proposed_lp = tgt.logp(proposed.data());                               // :309
  T log_q_forward = iso_proposal_logp(x, proposed.data(), d, prop_std);    // :310
  T log_q_backward = iso_proposal_logp(proposed.data(), x, d, prop_std);   // :311
  r.log_accept_ratio = (r.proposed_lp + log_q_backward) - (r.current_lp + log_q_forward);  // :312
  r.accepted = (r.log_accept_ratio > ln_u) ? 1 : 0;                        // :314 (strict)
  if (r.accepted) std::copy(proposed.begin(), proposed.end(), x);
  return r;
}

// ------------------------------------------------------------------------------------------
// Integer-state Metropolis-Hastings: MetropolisHastings<S = i32, T = f64> with the discrete targets and the
// +-1 random-walk proposals of /root/reference/tests/metrohast_poisson_test.rs:18-86 (Poisson) and :195-252
// (Binomial).  The proposal's direction bits are injected (step = +1 / -1 per coordinate).
// ------------------------------------------------------------------------------------------
enum IntTargetKind { IT_POISSON = 0, IT_BINOMIAL = 1 };

inline double ln_factorial(int k) {                 // metrohast_poisson_test.rs:40-50
  if (k < 2) return 0.0;
  double acc = 0.0;
  for (int i = 1; i <= k; ++i) acc += std::log((double)i);
  return acc;
}

struct IntTarget {
  int kind = IT_POISSON;
  int dim = 1;
  double lambda = 1.0, p = 0.5;
  int n = 0;
  // unnorm_logp: metrohast_poisson_test.rs:24-36 (Poisson), :203-213 (Binomial); coordinates are independent and
  // summed in order (the reference's examples are one-dimensional)
  double logp(const int* k) const {
    double total = 0.0;
    for (int i = 0; i < dim; ++i) {
      double lp;
      if (kind == IT_POISSON) {
        if (k[i] < 0) return -std::numeric_limits<double>::infinity();
        const double kf = (double)k[i];
        lp = kf * std::log(lambda) - lambda - ln_factorial(k[i]);
      } else {
        if (k[i] < 0 || k[i] > n) return -std::numeric_limits<double>::infinity();
        const double kf = (double)k[i], nf = (double)n;
        const double coeff = ln_factorial(n) - ln_factorial(k[i]) - ln_factorial(n - k[i]);
        lp = coeff + kf * std::log(p) + (nf - kf) * std::log(1.0 - p);
      }
      total = (i == 0) ? lp : total + lp;
    }
    return total;
  }
  // PoissonRandomWalk::sample :67-78 (reflect below 0 = clamp at 0), BinomialRandomWalk::sample :236-241 (clamp to [0, n])
  int propose(int cur, int step) const {
    const int v = cur + step;
    if (kind == IT_POISSON) return v < 0 ? 0 : v;
    return std::min(std::max(v, 0), n);
  }
};

struct MhIntStepInfo { double log_accept_ratio; int accepted; };

// MHMarkovChain::step, metropolis_hastings.rs:306-318, with Proposal::logp == ln(0.5) in both directions (:80-83)
inline MhIntStepInfo mh_int_step(const IntTarget& tgt, int* x, const signed char* steps, double ln_u) {
  std::vector<int> proposed(tgt.dim);
  for (int i = 0; i < tgt.dim; ++i) proposed[i] = tgt.propose(x[i], (int)steps[i]);
  const double current_lp = tgt.logp(x);
  const double proposed_lp = tgt.logp(proposed.data());
  const double log_q = std::log(0.5);
  MhIntStepInfo r;
  r.log_accept_ratio = (proposed_lp + log_q) - (current_lp + log_q);
  r.accepted = (r.log_accept_ratio > ln_u) ? 1 : 0;
  if (r.accepted) std::copy(proposed.begin(), proposed.end(), x);
  return r;
}

// ------------------------------------------------------------------------------------------
// Gibbs sweeps.  GibbsMarkovChain::step gibbs.rs:89-105 with the conditionals of the reference's own tests
// (gibbs.rs:177-245).  Randomness is injected: normals[i] / uniforms[i] are the draws coordinate i's conditional takes.
// ------------------------------------------------------------------------------------------
enum GibbsCondKind { GC_CONSTANT = 0, GC_MIXTURE_XZ = 1 };

struct GibbsConditional {
  int kind = GC_CONSTANT;
  double c = 0.0;                                         // ConstantConditional :177-186
  double mu0 = 0, sigma0 = 1, mu1 = 0, sigma1 = 1, pi0 = 0.5;   // MixtureConditional :188-197
  // MixtureConditional::normal_pdf :200-205
  static double normal_pdf(double x, double mu, double sigma) {
    const double var = sigma * sigma;
    const double coeff = 1.0 / std::sqrt(2.0 * 3.14159265358979323846 * var);
    const double dx = x - mu;
    const double exp_val = std::exp(-(dx * dx) / (2.0 * var));
    return coeff * exp_val;
  }
  // Conditional::sample :208-243
  double sample(int i, const double* given, double normal, double uniform) const {
    if (kind == GC_CONSTANT) return c;
    if (i == 0) {
      const double z = given[1];
      return z < 0.5 ? mu0 + sigma0 * normal : mu1 + sigma1 * normal;
    }
    const double x = given[0];
    const double p0 = pi0 * normal_pdf(x, mu0, sigma0);
    const double p1 = (1.0 - pi0) * normal_pdf(x, mu1, sigma1);
    const double total = p0 + p1;
    const double prob_z1 = total > 0.0 ? p1 / total : 0.5;
    return uniform < prob_z1 ? 1.0 : 0.0;
  }
};

// one full sweep, gibbs.rs:96-99: coordinate i sees the already-updated coordinates 0 .. i - 1
inline void gibbs_step(const GibbsConditional& cond, double* state, int dim, const double* normals, const double* uniforms) {
  for (int i = 0; i < dim; ++i) state[i] = cond.sample(i, state, normals[i], uniforms[i]);
}

// ------------------------------------------------------------------------------------------
// NUTS (identity mass matrix).  generic_nuts.rs:755-925, 1025-1102, 1153-1418
// ------------------------------------------------------------------------------------------
// Injected random stream, consumed in the reference's draw order:
//   per step: d normals -> Exp1 -> per doubling: u (direction) -> [post-order in the tree: one f64
//   uniform per internal node whose left subtree had s'] -> u (accept).
struct NutsStream {
  const double* normals = nullptr;  size_t n_normals = 0, i_normals = 0;
  const double* exp1 = nullptr;     size_t n_exp1 = 0, i_exp1 = 0;
  const double* unif = nullptr;     size_t n_unif = 0, i_unif = 0;
  bool exhausted = false;
  double next_normal() { if (i_normals >= n_normals) { exhausted = true; return 0.0; } return normals[i_normals++]; }
  double next_exp1()   { if (i_exp1 >= n_exp1) { exhausted = true; return 1.0; } return exp1[i_exp1++]; }
  double next_unif()   { if (i_unif >= n_unif) { exhausted = true; return 0.75; } return unif[i_unif++]; }
};

// MassMatrix::{Identity, Diagonal, Dense}, generic_nuts.rs:177-304.  kind 0 = Identity, 1 = Diagonal, 2 = Dense.
// cholesky_spd :306-331
template <class T>
inline bool cholesky_spd(const std::vector<T>& a, int dim, std::vector<T>& l) {
  l.assign((size_t)dim * dim, T(0));
  for (int i = 0; i < dim; ++i) {
    for (int j = 0; j <= i; ++j) {
      T sum = a[(size_t)i * dim + j];
      for (int k = 0; k < j; ++k) sum = sum - l[(size_t)i * dim + k] * l[(size_t)j * dim + k];
      if (i == j) {
        if (sum <= T(0) || !std::isfinite(sum)) return false;
        l[(size_t)i * dim + j] = std::sqrt(sum);
      } else {
        T dd = l[(size_t)j * dim + j];
        if (dd <= T(0) || !std::isfinite(dd)) return false;
        l[(size_t)i * dim + j] = sum / dd;
      }
    }
  }
  return true;
}
// invert_spd_from_cholesky :333-359
template <class T>
inline bool invert_spd_from_cholesky(const std::vector<T>& l, int dim, std::vector<T>& inv) {
  std::vector<T> inv_l((size_t)dim * dim, T(0));
  for (int i = 0; i < dim; ++i) {
    T dd = l[(size_t)i * dim + i];
    if (dd <= T(0) || !std::isfinite(dd)) return false;
    inv_l[(size_t)i * dim + i] = T(1) / dd;
    for (int j = i + 1; j < dim; ++j) {
      T sum = T(0);
      for (int k = i; k < j; ++k) sum = sum + l[(size_t)j * dim + k] * inv_l[(size_t)k * dim + i];
      inv_l[(size_t)j * dim + i] = -sum / l[(size_t)j * dim + j];
    }
  }
  inv.assign((size_t)dim * dim, T(0));
  for (int i = 0; i < dim; ++i) {
    for (int j = 0; j <= i; ++j) {
      T sum = T(0);
      for (int k = std::max(i, j); k < dim; ++k) sum = sum + inv_l[(size_t)k * dim + i] * inv_l[(size_t)k * dim + j];
      inv[(size_t)i * dim + j] = sum;
      inv[(size_t)j * dim + i] = sum;
    }
  }
  return true;
}

template <class T>
struct DiagMass {
  int kind = 0;                 // 0 Identity, 1 Diagonal, 2 Dense
  int dim = 0;
  std::vector<T> inv, sqrt_;    // Diagonal: [d], [d];  Dense: inv [d*d]
  std::vector<T> chol;          // Dense: lower Cholesky factor of the covariance [d*d]
  bool identity() const { return kind == 0; }
  // diagonal_from_var, generic_nuts.rs:196-206
  static DiagMass from_var(std::vector<T> var, T jitter) {
    DiagMass m;
    m.kind = 1; m.dim = (int)var.size();
    m.inv.resize(var.size()); m.sqrt_.resize(var.size());
    for (size_t i = 0; i < var.size(); ++i) {
      T v = std::max(var[i], jitter);
      m.inv[i] = T(1) / v;
      m.sqrt_[i] = std::sqrt(v);
    }
    return m;
  }
  // dense_from_cov, generic_nuts.rs:208-226: up to 8 tries with the diagonal jitter growing tenfold
  static bool dense_from_cov(const std::vector<T>& cov, int dim, T jitter, DiagMass* out) {
    T j = std::max(jitter, (T)1e-10);
    for (int t = 0; t < 8; ++t) {
      std::vector<T> cov_try = cov;
      for (int d = 0; d < dim; ++d) cov_try[(size_t)d * dim + d] = cov_try[(size_t)d * dim + d] + j;
      std::vector<T> chol, inv;
      if (cholesky_spd(cov_try, dim, chol) && invert_spd_from_cholesky(chol, dim, inv)) {
        out->kind = 2; out->dim = dim; out->inv = inv; out->chol = chol; out->sqrt_.clear();
        return true;
      }
      j = j * (T)10.0;
    }
    return false;
  }
  // inv_mul :265-281
  void inv_mul(const T* in, T* out, int d) const {
    if (kind == 0) { for (int i = 0; i < d; ++i) out[i] = in[i]; }
    else if (kind == 1) { for (int i = 0; i < d; ++i) out[i] = inv[i] * in[i]; }
    else {
      for (int i = 0; i < d; ++i) {
        T acc = T(0);
        for (int jj = 0; jj < d; ++jj) acc = acc + inv[(size_t)i * d + jj] * in[jj];
        out[i] = acc;
      }
    }
  }
  // sample_momentum :283-303 applied to already drawn standard normals
  void scale_momentum(T* z, int d) const {
    if (kind == 1) { for (int i = 0; i < d; ++i) z[i] = z[i] * sqrt_[i]; }
    else if (kind == 2) {
      std::vector<T> zz(z, z + d);
      for (int i = 0; i < d; ++i) {
        T acc = T(0);
        for (int jj = 0; jj <= i; ++jj) acc = acc + chol[(size_t)i * d + jj] * zz[jj];
        z[i] = acc;
      }
    }
  }
};

template <class T>
inline T nuts_kinetic(const T* p, int d, const DiagMass<T>* mass = nullptr) {  // generic_nuts.rs:228-263
  T q = 0;
  if (!mass || mass->kind == 0) for (int i = 0; i < d; ++i) q = q + p[i] * p[i];
  else if (mass->kind == 1) for (int i = 0; i < d; ++i) q = q + p[i] * p[i] * mass->inv[i];
  else {
    for (int i = 0; i < d; ++i) {
      T row_dot = T(0);
      for (int j = 0; j < d; ++j) row_dot = row_dot + mass->inv[(size_t)i * d + j] * p[j];
      q = q + p[i] * row_dot;
    }
  }
  return T(0.5) * q;
}

template <class T>
inline T nuts_leapfrog(const Target<T>& tgt, T* q, T* p, T* g, T eps, const DiagMass<T>* mass = nullptr) {  // generic_nuts.rs:1396-1418
  const int d = tgt.dim;
  const T half = T(0.5);
  add_scaled_assign(p, g, eps * half, d);
  if (!mass || mass->identity()) {
    add_scaled_assign(q, p, eps, d);  // identity mass: velocity = momentum
  } else {
    std::vector<T> vel(d);
    mass->inv_mul(p, vel.data(), d);   // apply_inv_mass
    add_scaled_assign(q, vel.data(), eps, d);
  }
  T logp = tgt.logp_and_grad(q, g);
  add_scaled_assign(p, g, eps * half, d);
  return logp;
}

template <class T>
inline bool nuts_stop_criterion(const T* qm, const T* qp, const T* pm, const T* pp, int d, const DiagMass<T>* mass = nullptr) {
  // generic_nuts.rs:1357-1378: diff = q+ - q- ; diff.v- >= 0 && diff.v+ >= 0, v = M^-1 p
  T dm = 0, dp = 0;
  if (!mass || mass->identity()) {
    for (int i = 0; i < d; ++i) { T df = qp[i] - qm[i]; dm = dm + df * pm[i]; }
    for (int i = 0; i < d; ++i) { T df = qp[i] - qm[i]; dp = dp + df * pp[i]; }
  } else {
    std::vector<T> vm(d), vp(d);
    mass->inv_mul(pm, vm.data(), d);
    mass->inv_mul(pp, vp.data(), d);
    for (int i = 0; i < d; ++i) { T df = qp[i] - qm[i]; dm = dm + df * vm[i]; }
    for (int i = 0; i < d; ++i) { T df = qp[i] - qm[i]; dp = dp + df * vp[i]; }
  }
  return dm >= T(0) && dp >= T(0);
}

template <class T>
struct TreeOut {
  std::vector<T> q_minus, p_minus, g_minus, q_plus, p_plus, g_plus, q_prime, g_prime;
  T logp_prime = 0;
  size_t n_prime = 0;
  bool s_prime = false;
  T alpha_prime = 0;
  size_t n_alpha_prime = 0;
};

// build_tree_with_mass, generic_nuts.rs:1153-1341 (recursive, identity mass).  `leapfrogs` counts
// gradient evaluations; `max_leaves_guard` is not in the reference (safety for tests only).
template <class T>
TreeOut<T> nuts_build_tree(const Target<T>& tgt, const std::vector<T>& q, const std::vector<T>& p,
                           const std::vector<T>& g, T logu, int v, int j, T eps, T joint_0,
                           NutsStream& rng, size_t* leapfrogs, const DiagMass<T>* mass = nullptr) {
  const int d = tgt.dim;
  if (j == 0) {
    TreeOut<T> o;
    std::vector<T> q1 = q, p1 = p, g1 = g;
    T logp1 = nuts_leapfrog(tgt, q1.data(), p1.data(), g1.data(), (T)v * eps, mass);  // :1185-1192
    if (leapfrogs) ++*leapfrogs;
    T joint = logp1 - nuts_kinetic(p1.data(), d, mass);                         // :1193
    o.n_prime = (logu < joint) ? 1 : 0;                                         // :1194
    o.s_prime = (logu - T(1000)) < joint;                                       // :1195
    o.q_minus = q1; o.q_plus = q1; o.p_minus = p1; o.p_plus = p1; o.g_minus = g1; o.g_plus = g1;
    o.q_prime = q1; o.g_prime = g1; o.logp_prime = logp1;
    o.alpha_prime = std::min(T(1), (T)std::exp(joint - joint_0));               // :1202
    o.n_alpha_prime = 1;
    return o;
  }
  TreeOut<T> o = nuts_build_tree(tgt, q, p, g, logu, v, j - 1, eps, joint_0, rng, leapfrogs, mass);  // :1238
  if (o.s_prime) {                                                                              // :1251
    TreeOut<T> o2 = (v == -1)
        ? nuts_build_tree(tgt, o.q_minus, o.p_minus, o.g_minus, logu, v, j - 1, eps, joint_0, rng, leapfrogs, mass)
        : nuts_build_tree(tgt, o.q_plus, o.p_plus, o.g_plus, logu, v, j - 1, eps, joint_0, rng, leapfrogs, mass);
    if (v == -1) { o.q_minus = o2.q_minus; o.p_minus = o2.p_minus; o.g_minus = o2.g_minus; }
    else         { o.q_plus = o2.q_plus;   o.p_plus = o2.p_plus;   o.g_plus = o2.g_plus; }
    double u_build_tree = rng.next_unif();                                                      // :1305 (f64)
    size_t den = std::max<size_t>(o.n_prime + o2.n_prime, 1);
    if (u_build_tree < ((double)o2.n_prime / (double)den)) {                                    // :1306
      o.q_prime = o2.q_prime; o.g_prime = o2.g_prime; o.logp_prime = o2.logp_prime;
    }
    o.n_prime += o2.n_prime;                                                                    // :1312
    o.s_prime = o.s_prime && o2.s_prime &&
                nuts_stop_criterion(o.q_minus.data(), o.q_plus.data(), o.p_minus.data(), o.p_plus.data(), d);  // :1314-1321
    o.alpha_prime = o.alpha_prime + o2.alpha_prime;                                             // :1322
    o.n_alpha_prime += o2.n_alpha_prime;
  }
  return o;
}

template <class T>
inline bool all_finite(const T* v, int d) {
  for (int i = 0; i < d; ++i) if (!std::isfinite(v[i])) return false;
  return true;
}

// find_reasonable_epsilon_with_mass, generic_nuts.rs:1025-1102.  Both reference call sites (:745, :911) go through
// find_reasonable_epsilon (:1009-1023), which passes the IDENTITY mass even when the chain carries an adapted one.
template <class T>
T nuts_find_reasonable_epsilon(const Target<T>& tgt, const T* position, const T* mom, const DiagMass<T>* mass = nullptr) {
  const int d = tgt.dim;
  T epsilon = 1;
  const T half = T(0.5);
  std::vector<T> grad(d, T(0));
  T ulogp = tgt.logp_and_grad(position, grad.data());
  std::vector<T> q1(position, position + d), p1(mom, mom + d), g1 = grad;
  T ulogp1 = nuts_leapfrog(tgt, q1.data(), p1.data(), g1.data(), epsilon, mass);
  T k = 1;
  while (!std::isfinite(ulogp1) || !all_finite(g1.data(), d)) {
    k = k * half;
    q1.assign(position, position + d); p1.assign(mom, mom + d); g1 = grad;
    ulogp1 = nuts_leapfrog(tgt, q1.data(), p1.data(), g1.data(), epsilon * k, mass);
  }
  epsilon = half * k * epsilon;
  T lap = ulogp1 - ulogp - (nuts_kinetic(p1.data(), d, mass) - nuts_kinetic(mom, d, mass));
  T a = (lap > std::log(half)) ? T(1) : T(-1);
  while (a * lap > -a * std::log(T(2))) {
    epsilon = epsilon * std::pow(T(2), a);
    q1.assign(position, position + d); p1.assign(mom, mom + d); g1 = grad;
    ulogp1 = nuts_leapfrog(tgt, q1.data(), p1.data(), g1.data(), epsilon, mass);
    lap = ulogp1 - ulogp - (nuts_kinetic(p1.data(), d, mass) - nuts_kinetic(mom, d, mass));
  }
  return epsilon;
}

// GenericNUTSChain state + step (generic_nuts.rs:560-581, 630-647, 731-753, 755-925)
template <class T>
struct NutsChain {
  Target<T> tgt;
  std::vector<T> position;
  T target_accept_p = T(0.8);
  T epsilon = T(-1);
  size_t m = 0, n_collect = 0, n_discard = 0;
  T gamma = T(0.05);
  size_t t_0 = 10;
  T kappa = T(0.75);
  T mu = std::log(T(10));
  T epsilon_bar = T(1);
  T h_bar = T(0);
  int max_depth = 0;  // 0 = uncapped (reference); >0 = cap added by BASELINE cfg5
  // diagonal mass-matrix adaptation (NUTSMassMatrixConfig / MassMatrixWarmup / RunningCov, generic_nuts.rs:40-175)
  bool mass_adapt = false;
  size_t start_buffer = 75, end_buffer = 50, initial_window = 25;
  double regularize = 0.05, jitter = 1e-6;
  DiagMass<T> mass;
  size_t next_window_end = 0, window_len = 0, run_n = 0;
  std::vector<T> run_mean, run_m2;
  // dense adaptation (MassMatrixAdaptation::Dense): config_dense = what the config asks for; running_dense = whether the
  // running covariance keeps the full matrix (new_shared falls back to diagonal statistics when dim > dense_max_dim, :612-628,
  // while maybe_update_mass_matrix still dispatches on the CONFIG, :962-996: with the fallback the mass is never updated)
  bool config_dense = false, running_dense = false;
  std::vector<T> run_m2_dense;
  size_t mass_updates = 0;
  void enable_mass_adaptation(size_t sb, size_t eb, size_t iw, double reg, double jit, bool dense = false, size_t dense_max_dim = 75) {
    mass_adapt = true; start_buffer = sb; end_buffer = eb; initial_window = iw; regularize = reg; jitter = jit;
    size_t sbm = std::max<size_t>(sb, 1);                 // MassMatrixWarmup::new, :141-150
    window_len = std::max<size_t>(iw, 10);
    next_window_end = sbm + window_len;
    run_mean.assign(tgt.dim, T(0)); run_m2.assign(tgt.dim, T(0)); run_n = 0;
    config_dense = dense;
    running_dense = dense && (size_t)tgt.dim <= dense_max_dim;
    if (running_dense) run_m2_dense.assign((size_t)tgt.dim * tgt.dim, T(0));
  }
  bool should_collect(size_t mm, size_t n_warm) const {   // :152-160
    if (mm == 0 || mm > n_warm) return false;
    if (mm <= start_buffer) return false;
    size_t lim = n_warm > end_buffer ? n_warm - end_buffer : 0;
    return mm < lim;
  }
  bool note_if_window_end(size_t mm, size_t n_warm) {     // :162-173
    if (!should_collect(mm, n_warm)) return false;
    size_t lim = n_warm > end_buffer ? n_warm - end_buffer : 0;
    if (mm >= next_window_end || mm + 1 >= lim) {
      next_window_end = next_window_end + window_len;
      window_len = std::min<size_t>(window_len * 2, 400);
      return true;
    }
    return false;
  }
  // diagnostics of the last step
  size_t last_leapfrogs = 0;
  int last_depth = 0;

  // init_chain_state, generic_nuts.rs:731-753 (consumes d normals)
  void init_chain_state(size_t n_collect_, size_t n_discard_, NutsStream& rng) {
    const int d = tgt.dim;
    n_collect = n_collect_; n_discard = n_discard_; m = 0;
    std::vector<T> mom0(d);
    for (int i = 0; i < d; ++i) mom0[i] = (T)rng.next_normal();
    mass.scale_momentum(mom0.data(), d);                                                     // sample_momentum :283-303
    if (mass_adapt) {
      run_n = 0; std::fill(run_mean.begin(), run_mean.end(), T(0)); std::fill(run_m2.begin(), run_m2.end(), T(0));
      std::fill(run_m2_dense.begin(), run_m2_dense.end(), T(0));
    }
    if (std::abs(epsilon + T(1)) <= std::numeric_limits<T>::epsilon())
      epsilon = nuts_find_reasonable_epsilon(tgt, position.data(), mom0.data());   // :745 — identity mass (find_reasonable_epsilon :1009-1023)
    mu = std::log(T(10) * epsilon);
  }

  void step(NutsStream& rng) {
    const int d = tgt.dim;
    m += 1;
    std::vector<T> mom0(d);
    for (int i = 0; i < d; ++i) mom0[i] = (T)rng.next_normal();                    // :761
    mass.scale_momentum(mom0.data(), d);
    std::vector<T> grad(d, T(0));
    T logp = tgt.logp_and_grad(position.data(), grad.data());                     // :765
    T joint = logp - nuts_kinetic(mom0.data(), d, &mass);                         // :766
    T exp1_obs = (T)rng.next_exp1();                                              // :767
    T logu = joint - exp1_obs;                                                    // :768
    std::vector<T> qm = position, qp = position, pm = mom0, pp = mom0, gm = grad, gp = grad;
    int j = 0;
    size_t n = 1;
    bool s = true;
    T alpha = 0;
    size_t n_alpha = 0;
    size_t leap = 0;
    while (s) {                                                                   // :782
      T u_run_1 = (T)rng.next_unif();
      int v = (u_run_1 < T(0.5)) ? 1 : -1;                                        // :784
      TreeOut<T> o = (v == -1)
          ? nuts_build_tree(tgt, qm, pm, gm, logu, v, j, epsilon, joint, rng, &leap, &mass)
          : nuts_build_tree(tgt, qp, pp, gp, logu, v, j, epsilon, joint, rng, &leap, &mass);
      if (v == -1) { qm = o.q_minus; pm = o.p_minus; gm = o.g_minus; }
      else         { qp = o.q_plus;  pp = o.p_plus;  gp = o.g_plus; }
      alpha = o.alpha_prime; n_alpha = o.n_alpha_prime;
      T tmp = std::min(T(1), (T)o.n_prime / (T)n);                                // :860-864
      T u_run_2 = (T)rng.next_unif();                                             // :865
      if (o.s_prime && (u_run_2 < tmp)) position = o.q_prime;                     // :866-868
      n += o.n_prime;                                                             // :869
      s = o.s_prime && nuts_stop_criterion(qm.data(), qp.data(), pm.data(), pp.data(), d, &mass);  // :871-878
      j += 1;
      if (max_depth > 0 && j >= max_depth) s = false;  // cap: not in the reference
      if (rng.exhausted) break;
    }
    last_leapfrogs = leap; last_depth = j;
    // dual averaging, generic_nuts.rs:882-924
    T eta = T(1) / (T)(m + t_0);
    h_bar = (T(1) - eta) * h_bar + eta * (target_accept_p - alpha / (T)n_alpha);
    if (m <= n_discard) {
      T mm = (T)m;
      epsilon = std::exp(mu - std::sqrt(mm) / gamma * h_bar);
      eta = std::pow(mm, -kappa);
      epsilon_bar = std::exp((T(1) - eta) * std::log(epsilon_bar) + eta * std::log(epsilon));
      if (mass_adapt && should_collect(m, n_discard)) {                             // :902-920
        // RunningCov::update, :105-126
        run_n += 1;
        T n_s = (T)run_n;
        std::vector<T> delta(d);
        for (int i = 0; i < d; ++i) {
          delta[i] = position[i] - run_mean[i];
          run_mean[i] = run_mean[i] + delta[i] / n_s;
          T delta2 = position[i] - run_mean[i];
          run_m2[i] = run_m2[i] + delta[i] * delta2;
        }
        if (running_dense) {
          std::vector<T> delta2(d);
          for (int i = 0; i < d; ++i) delta2[i] = position[i] - run_mean[i];
          for (int i = 0; i < d; ++i)
            for (int j = i; j < d; ++j) run_m2_dense[(size_t)i * d + j] = run_m2_dense[(size_t)i * d + j] + delta[i] * delta2[j];
        }
        if (note_if_window_end(m, n_discard) && run_n >= 5) {                       // maybe_update_mass_matrix :948-997
          T n_denom = (T)(run_n - 1);
          T reg = (T)regularize, omr = T(1) - reg;
          T jit = (T)std::max(jitter, 1e-10);
          bool updated = false;
          if (!config_dense) {
            std::vector<T> var(d);
            for (int i = 0; i < d; ++i) var[i] = std::max(omr * (run_m2[i] / n_denom) + reg, jit);
            mass = DiagMass<T>::from_var(var, jit);
            updated = true;
          } else if (running_dense) {
            std::vector<T> cov((size_t)d * d, T(0));
            for (int i = 0; i < d; ++i)
              for (int j = i; j < d; ++j) {
                T raw = run_m2_dense[(size_t)i * d + j] / n_denom;
                T v = (i == j) ? std::max(omr * raw + reg, jit) : omr * raw;
                cov[(size_t)i * d + j] = v; cov[(size_t)j * d + i] = v;
              }
            DiagMass<T> nm;
            if (DiagMass<T>::dense_from_cov(cov, d, jit, &nm)) { mass = nm; updated = true; }
            else if (mass.kind == 0) { mass = DiagMass<T>::from_var(std::vector<T>(d, T(1)), jit); updated = true; }   // :989-994
          }                                                     // config Dense with diagonal statistics: None (:972-974)
          if (updated) {
            mass_updates += 1;
            std::vector<T> probe(d);
            for (int i = 0; i < d; ++i) probe[i] = (T)rng.next_normal();
            mass.scale_momentum(probe.data(), d);
            epsilon = nuts_find_reasonable_epsilon(tgt, position.data(), probe.data());   // :911 — identity mass, as the reference
            mu = std::log(T(10) * epsilon);
            epsilon_bar = epsilon;
            h_bar = T(0);
            run_n = 0; std::fill(run_mean.begin(), run_mean.end(), T(0)); std::fill(run_m2.begin(), run_m2.end(), T(0));
            std::fill(run_m2_dense.begin(), run_m2_dense.end(), T(0));
          }
        }
      }
    } else {
      epsilon = epsilon_bar;
    }
  }
};

// ------------------------------------------------------------------------------------------
// Diagnostics (all f32, as the reference).  stats.rs:342-368, 419-681, 199-339
// ------------------------------------------------------------------------------------------
// splitcat, stats.rs:419-425: [c,n,p] -> [2c, n/2, p]; second half = the LAST n/2 draws.
inline std::vector<float> splitcat(const float* s, size_t c, size_t n, size_t p, size_t* out_c, size_t* out_n) {
  size_t half = n / 2;
  std::vector<float> out(2 * c * half * p);
  for (size_t ch = 0; ch < c; ++ch)
    for (size_t t = 0; t < half; ++t)
      for (size_t k = 0; k < p; ++k) {
        out[(ch * half + t) * p + k] = s[(ch * n + t) * p + k];
        out[((c + ch) * half + t) * p + k] = s[(ch * n + (n - half) + t) * p + k];
      }
  *out_c = 2 * c; *out_n = half;
  return out;
}

// withinvar, stats.rs:456-504 (here c,n are the split shape)
inline void withinvar(const float* s, size_t c, size_t n, size_t p, float* within, float* var) {
  for (size_t k = 0; k < p; ++k) {
    std::vector<float> cm(c);
    for (size_t ch = 0; ch < c; ++ch) {
      float sum = 0.f;
      for (size_t t = 0; t < n; ++t) sum += s[(ch * n + t) * p + k];
      cm[ch] = sum / (float)n;
    }
    float om = 0.f;
    for (size_t ch = 0; ch < c; ++ch) om += cm[ch];
    om = om / (float)c;
    float ss = 0.f;
    for (size_t ch = 0; ch < c; ++ch) { float df = cm[ch] - om; ss += df * df; }
    float b = ss * ((float)n / (float)(c - 1));
    float wsum = 0.f;
    for (size_t ch = 0; ch < c; ++ch) {
      float sq = 0.f;
      for (size_t t = 0; t < n; ++t) { float v = s[(ch * n + t) * p + k]; sq += (v - cm[ch]) * (v - cm[ch]); }
      wsum += sq / (float)n;
    }
    float w = wsum / (float)c;
    float v = (((float)n - 1.0f) / (float)n) * w + b / (float)n;
    within[k] = w; var[k] = v;
  }
}

// autocov_bf, stats.rs:659-681: sample [n,d] -> out [n,d]
inline void autocov_bf(const float* x, size_t n, size_t d, float* out) {
  std::vector<float> col(n);
  for (size_t k = 0; k < d; ++k) {
    float sum = 0.f;
    for (size_t t = 0; t < n; ++t) sum += x[t * d + k];
    float mean = sum / (float)n;
    for (size_t t = 0; t < n; ++t) col[t] = x[t * d + k] - mean;
    for (size_t lag = 0; lag < n; ++lag) {
      float sl = 0.f;
      for (size_t t = 0; t < n - lag; ++t) sl += col[t] * col[t + lag];
      out[lag * d + k] = sl / (float)n;
    }
  }
}

// radix-2 complex FFT in f32 (stand-in for rustfft 6.4, absent from /root/reference; pinned to
// 1e-6 by stats.rs:808-839)
inline void fft_inplace(std::vector<float>& re, std::vector<float>& im, bool inverse) {
  size_t n = re.size();
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    double ang = 2.0 * 3.14159265358979323846 / (double)len * (inverse ? 1.0 : -1.0);
    for (size_t i = 0; i < n; i += len)
      for (size_t k = 0; k < len / 2; ++k) {
        float wr = (float)std::cos(ang * (double)k), wi = (float)std::sin(ang * (double)k);
        size_t a = i + k, b = i + k + len / 2;
        float xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] = re[a] + xr; im[a] = im[a] + xi;
      }
  }
}

// autocov_fft, stats.rs:603-647
inline void autocov_fft(const float* x, size_t n, size_t d, float* out) {
  size_t n_padded = 1;
  while (n_padded < 2 * n - 1) n_padded <<= 1;
  std::vector<float> re(n_padded), im(n_padded);
  for (size_t k = 0; k < d; ++k) {
    float sum = 0.f;
    for (size_t t = 0; t < n; ++t) sum += x[t * d + k];
    float mean = sum / (float)n;
    for (size_t t = 0; t < n_padded; ++t) { re[t] = t < n ? x[t * d + k] - mean : 0.f; im[t] = 0.f; }
    fft_inplace(re, im, false);
    for (size_t t = 0; t < n_padded; ++t) { re[t] = re[t] * re[t] + im[t] * im[t]; im[t] = 0.f; }
    fft_inplace(re, im, true);
    for (size_t t = 0; t < n; ++t) out[t * d + k] = re[t] / (float)n_padded / (float)n;
  }
}

inline void autocov(const float* x, size_t n, size_t d, float* out) {  // stats.rs:575-581
  if (n <= 100) autocov_bf(x, n, d, out); else autocov_fft(x, n, d, out);
}

// ess, stats.rs:523-573 (input: split sample [c,n,p])
inline void ess(const float* s, size_t c, size_t n, size_t p, const float* within, const float* var, float* out) {
  std::vector<float> avg(n * p, 0.f), ac(n * p);
  for (size_t ch = 0; ch < c; ++ch) {
    autocov(s + ch * n * p, n, p, ac.data());
    for (size_t i = 0; i < n * p; ++i) avg[i] += ac[i];
  }
  for (size_t i = 0; i < n * p; ++i) avg[i] = avg[i] / (float)c;
  for (size_t k = 0; k < p; ++k) {
    std::vector<float> rho(n);
    for (size_t t = 0; t < n; ++t) {
      float diff = -avg[t * p + k] + within[k];
      rho[t] = -(diff / var[k]) + 1.0f;
    }
    float mn = n >= 2 ? rho[0] + rho[1] : 0.0f;
    float acc = 0.0f;
    for (size_t t = 0; t + 1 < n; t += 2) {  // windows_with_stride(2,2)
      float p_t = rho[t] + rho[t + 1];
      if (p_t <= 0.0f) break;
      if (p_t > mn) p_t = mn;
      mn = p_t;
      acc += p_t;
    }
    float tau = -1.0f + 2.0f * acc;
    out[k] = (1.0f / tau) * (float)c * (float)n;
  }
}

// split_rhat_mean_ess, stats.rs:439-454.  NOTE rhat = sqrt(within / var) (inverted vs Stan).
inline void split_rhat_mean_ess(const float* sample, size_t c, size_t n, size_t p, float* rhat_out, float* ess_out) {
  size_t c2, n2;
  std::vector<float> sp = splitcat(sample, c, n, p, &c2, &n2);
  std::vector<float> within(p), var(p);
  withinvar(sp.data(), c2, n2, p, within.data(), var.data());
  for (size_t k = 0; k < p; ++k) rhat_out[k] = std::sqrt(within[k] / var[k]);
  ess(sp.data(), c2, n2, p, within.data(), var.data(), ess_out);
}

struct BasicStats { float min, median, max, mean, std; };
// basic_stats, stats.rs:342-368 (sort descending; median = data[len/2]; std ddof=1)
inline BasicStats basic_stats(std::vector<float> data) {
  std::sort(data.begin(), data.end(), [](float a, float b) { return a > b; });
  BasicStats r;
  r.min = data.back(); r.median = data[data.size() / 2]; r.max = data.front();
  float sum = 0.f;
  for (float v : data) sum += v;
  r.mean = sum / (float)data.size();
  float ss = 0.f;
  for (float v : data) ss += (v - r.mean) * (v - r.mean);
  r.std = std::sqrt(ss / ((float)data.size() - 1.0f));
  return r;
}

// MultiChainTracker, stats.rs:199-339 (running mean / mean_sq, EMA p_accept, rhat = sqrt(var/within))
struct MultiChainTracker {
  size_t n = 0, n_chains, n_params;
  float p_accept = 0.f;
  std::vector<float> last_state, mean, mean_sq;
  MultiChainTracker(size_t c, size_t p) : n_chains(c), n_params(p), last_state(c * p, 0.f), mean(c * p, 0.f), mean_sq(c * p, 0.f) {}
  void step(const float* x) {  // stats.rs:240-268
    n += 1;
    float nf = (float)n;
    for (size_t i = 0; i < n_chains * n_params; ++i) {
      mean[i] = (mean[i] * (nf - 1.0f) + x[i]) / nf;
      if (n == 1) mean_sq[i] = x[i] * x[i];
      else mean_sq[i] = (mean_sq[i] * (nf - 1.0f) + x[i] * x[i]) / nf;
    }
    for (size_t ch = 0; ch < n_chains; ++ch) {
      bool ne = false;
      for (size_t k = 0; k < n_params; ++k) ne = ne || (x[ch * n_params + k] != last_state[ch * n_params + k]);
      p_accept = (1.0f - 0.01f) * p_accept + 0.01f * (ne ? 1.0f : 0.0f);
    }
    std::copy(x, x + n_chains * n_params, last_state.begin());
  }
  void rhat(float* out) const {  // stats.rs:314-339
    float nc = (float)n_chains, nf = (float)n;
    float fac = nf / (nc - 1.0f);
    for (size_t k = 0; k < n_params; ++k) {
      float mc = 0.f;
      for (size_t ch = 0; ch < n_chains; ++ch) mc += mean[ch * n_params + k];
      mc = mc / nc;
      float between = 0.f, wsum = 0.f;
      for (size_t ch = 0; ch < n_chains; ++ch) {
        float df = mean[ch * n_params + k] - mc;
        between += df * df;
        float m1 = mean[ch * n_params + k];
        wsum += (mean_sq[ch * n_params + k] - m1 * m1) * nf / (nf - 1.0f);
      }
      between = between * fac;
      float within = wsum / nc;
      float var = within * ((nf - 1.0f) / nf) + between * (1.0f / nf);
      out[k] = std::sqrt(var / within);
    }
  }
};

// ------------------------------------------------------------------------------------------
// Host Philox4x32-10 + Box–Muller.  NOT from the reference: it is the restatement of the
// product's counter-based stream (include/gmcmc.h "RNG contract"), used (i) to check the device
// integer stream bit-for-bit and (ii) to drive the timed CPU baseline without injected arrays.
// Algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Random123.
// ------------------------------------------------------------------------------------------
struct Philox {
  static inline void round1(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
  static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
      round1(c, k);
      k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
  }
};

// uniform in (0,1]: (x + 1) * 2^-32 for f32 (24-bit rounding may give exactly 1.0, never 0);
// 53-bit for f64.
inline float u01_f32(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
inline double u01_f64(uint32_t hi, uint32_t lo) {
  uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)v + 1.0) * (1.0 / 9007199254740992.0);
}

}  // namespace orc
