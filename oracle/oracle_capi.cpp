// oracle_capi.cpp — extern "C" surface of the CPU oracle for ctypes (tests/, smoke(), bench cpu_baseline).
// TEST INFRASTRUCTURE ONLY — see gmcmc_oracle.hpp header.  Build: oracle/Makefile.
#include "gmcmc_oracle.hpp"

#include <chrono>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {

template <class T>
T target_logp_grad(int kind, int dim, const double* params, size_t np, const T* x, T* g) {
  Target<T> t(kind, dim, params, np);
  return t.logp_and_grad(x, g);
}

template <class T>
T target_logp(int kind, int dim, const double* params, size_t np, const T* x) {
  Target<T> t(kind, dim, params, np);
  return t.logp(x);
}

// Runs n_steps HMC transitions for C chains with injected momenta [n_steps,C,d] and ln_u [n_steps,C].
// q [C,d] in/out.  Optional outputs: samples [C,n_steps,d], accepted [n_steps,C], log_accept [n_steps,C],
// prop_q/prop_p [n_steps,C,d] (end of trajectory before the accept test).
template <class T>
void hmc_run(int kind, int dim, const double* params, size_t np, size_t C, T* q, T eps, int L,
             size_t n_steps, const T* momenta, const T* ln_u, T* samples, uint8_t* accepted,
             T* log_accept, T* prop_q, T* prop_p, T* logp_cur, T* logp_prop) {
  Target<T> t(kind, dim, params, np);
  const size_t d = (size_t)dim;
  for (size_t s = 0; s < n_steps; ++s) {
#pragma omp parallel for schedule(static)
    for (long long ci = 0; ci < (long long)C; ++ci) {
      size_t c = (size_t)ci;
      HmcStepInfo<T> r = hmc_step(t, q + c * d, momenta + (s * C + c) * d, ln_u[s * C + c], eps, L,
                                  prop_q ? prop_q + (s * C + c) * d : nullptr,
                                  prop_p ? prop_p + (s * C + c) * d : nullptr);
      if (samples) std::memcpy(samples + (c * n_steps + s) * d, q + c * d, d * sizeof(T));
      if (accepted) accepted[s * C + c] = (uint8_t)r.accepted;
      if (log_accept) log_accept[s * C + c] = r.log_accept;
      if (logp_cur) logp_cur[s * C + c] = r.logp_current;
      if (logp_prop) logp_prop[s * C + c] = r.logp_proposed;
    }
  }
}

// MH with injected normals [n_steps,C,d] and ln_u [n_steps,C].  x [C,d] in/out; samples f64 [C,n_steps,d]
// (Trace -> f64, core.rs:34-51; layout core.rs:219-229).
template <class T>
void mh_run(int kind, int dim, const double* params, size_t np, T prop_std, size_t C, T* x,
            size_t n_steps, const T* normals, const T* ln_u, double* samples, uint8_t* accepted,
            T* log_ratio) {
  Target<T> t(kind, dim, params, np);
  const size_t d = (size_t)dim;
  for (size_t s = 0; s < n_steps; ++s) {
#pragma omp parallel for schedule(static)
    for (long long ci = 0; ci < (long long)C; ++ci) {
      size_t c = (size_t)ci;
      MhStepInfo<T> r = mh_step(t, prop_std, x + c * d, normals + (s * C + c) * d, ln_u[s * C + c]);
      if (samples)
        for (size_t k = 0; k < d; ++k) samples[(c * n_steps + s) * d + k] = (double)x[c * d + k];
      if (accepted) accepted[s * C + c] = (uint8_t)r.accepted;
      if (log_ratio) log_ratio[s * C + c] = r.log_accept_ratio;
    }
  }
}

template <class T>
void to_vec(std::vector<T>& v, const T* p, int d) { v.assign(p, p + d); }

template <class T>
void build_tree_c(int kind, int dim, const double* params, size_t np, const T* q, const T* p,
                  const T* g, T logu, int v, int j, T eps, T joint_0, const double* unif,
                  size_t n_unif, T* out_vecs /*8*d: q-,p-,g-,q+,p+,g+,q',g'*/,
                  T* out_scalars /*logp', alpha'*/, long long* out_ints /*n', s', n_alpha, leapfrogs, unif_used*/) {
  Target<T> t(kind, dim, params, np);
  NutsStream rng; rng.unif = unif; rng.n_unif = n_unif;
  std::vector<T> qv(q, q + dim), pv(p, p + dim), gv(g, g + dim);
  size_t leap = 0;
  TreeOut<T> o = nuts_build_tree(t, qv, pv, gv, logu, v, j, eps, joint_0, rng, &leap);
  const std::vector<T>* vs[8] = {&o.q_minus, &o.p_minus, &o.g_minus, &o.q_plus, &o.p_plus, &o.g_plus, &o.q_prime, &o.g_prime};
  for (int i = 0; i < 8; ++i) std::memcpy(out_vecs + (size_t)i * dim, vs[i]->data(), sizeof(T) * dim);
  out_scalars[0] = o.logp_prime; out_scalars[1] = o.alpha_prime;
  out_ints[0] = (long long)o.n_prime; out_ints[1] = o.s_prime ? 1 : 0; out_ints[2] = (long long)o.n_alpha_prime;
  out_ints[3] = (long long)leap; out_ints[4] = (long long)rng.i_unif;
}

// NUTS run for C chains with per-chain injected streams.
// normals [C, n_norm], exp1 [C, n_exp], unif [C, n_unif] (doubles).  Mirrors NUTS::run (nuts.rs:214-257):
// init_chain_state, then total = n_collect+n_discard iterations where iteration 0 takes no step;
// sample index step_idx - n_discard.  Outputs: samples [C,n_collect,d], eps_final [C], leapfrogs [C]
// (total gradient evaluations inside trees), used counts [C,3], depth_sum [C].
template <class T>
void nuts_run(int kind, int dim, const double* params, size_t np, size_t C, T* q, T target_accept,
              int max_depth, T eps_init /* <0 => find_reasonable_epsilon */, size_t n_collect,
              size_t n_discard, const double* normals, size_t n_norm, const double* exp1, size_t n_exp,
              const double* unif, size_t n_unif, T* samples, T* eps_final, long long* leapfrogs,
              long long* used, int* exhausted, const double* mass_cfg /*null or [start_buffer, end_buffer, initial_window,
              regularize, jitter, dense (0 / 1), dense_max_dim]*/ = nullptr, T* mass_inv_out /*[C,d] (diagonal) / [C,d,d] (dense) or null*/ = nullptr,
              int n_runs = 1, long long* mass_updates_out = nullptr) {
  const size_t d = (size_t)dim;
#pragma omp parallel for schedule(dynamic, 1)
  for (long long ci = 0; ci < (long long)C; ++ci) {
    size_t c = (size_t)ci;
    NutsChain<T> ch;
    ch.tgt = Target<T>(kind, dim, params, np);
    ch.position.assign(q + c * d, q + (c + 1) * d);
    ch.target_accept_p = target_accept;
    ch.max_depth = max_depth;
    if (eps_init > T(0)) ch.epsilon = eps_init;
    NutsStream rng;
    rng.normals = normals + c * n_norm; rng.n_normals = n_norm;
    rng.exp1 = exp1 + c * n_exp; rng.n_exp1 = n_exp;
    rng.unif = unif + c * n_unif; rng.n_unif = n_unif;
    const bool want_dense = mass_cfg && mass_cfg[5] != 0.0;
    if (mass_cfg) ch.enable_mass_adaptation((size_t)mass_cfg[0], (size_t)mass_cfg[1], (size_t)mass_cfg[2], mass_cfg[3], mass_cfg[4],
                                            want_dense, (size_t)mass_cfg[6]);
    long long leap = 0;
    for (int run = 0; run < n_runs; ++run) {          // n_runs > 1: repeated run() calls on the same chain (state carries over)
      ch.init_chain_state(n_collect, n_discard, rng);
      size_t total = n_collect + n_discard;
      for (size_t s = 0; s < total; ++s) {
        if (s > 0) { ch.step(rng); leap += (long long)ch.last_leapfrogs; }
        if (s >= n_discard && samples)
          std::memcpy(samples + (c * n_collect + (s - n_discard)) * d, ch.position.data(), d * sizeof(T));
      }
    }
    if (mass_inv_out) {
      if (want_dense) {       // [C, d, d]: identity until the first update, diag(inv) for the all-ones fallback
        for (size_t i = 0; i < d; ++i)
          for (size_t j = 0; j < d; ++j)
            mass_inv_out[(c * d + i) * d + j] = ch.mass.kind == 2 ? ch.mass.inv[i * d + j] : (i == j ? (ch.mass.kind == 1 ? ch.mass.inv[i] : T(1)) : T(0));
      } else {
        for (size_t i = 0; i < d; ++i) mass_inv_out[c * d + i] = ch.mass.identity() ? T(1) : ch.mass.inv[i];
      }
    }
    if (mass_updates_out) mass_updates_out[c] = (long long)ch.mass_updates;
    std::memcpy(q + c * d, ch.position.data(), d * sizeof(T));
    if (eps_final) eps_final[c] = ch.epsilon;
    if (leapfrogs) leapfrogs[c] = leap;
    if (used) { used[c * 3 + 0] = (long long)rng.i_normals; used[c * 3 + 1] = (long long)rng.i_exp1; used[c * 3 + 2] = (long long)rng.i_unif; }
    if (exhausted) exhausted[c] = rng.exhausted ? 1 : 0;
  }
}

// ---- timed CPU baselines (self-driven Philox randomness; "port" of the reference algorithm) ----
template <class T>
inline void philox_normals(uint64_t seed, uint64_t chain, uint32_t step, uint32_t stream, int d, T* out) {
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int i = 0; i < d; i += 4) {
    uint32_t ctr[4] = {(uint32_t)chain, (uint32_t)(chain >> 32), step, (stream << 24) | (uint32_t)(i / 4)};
    uint32_t r[4];
    Philox::block(ctr, key, r);
    float u0 = u01_f32(r[0]), u1 = u01_f32(r[1]), u2 = u01_f32(r[2]), u3 = u01_f32(r[3]);
    float r0 = std::sqrt(-2.0f * std::log(u0)), r1 = std::sqrt(-2.0f * std::log(u2));
    float z[4] = {r0 * std::cos(6.28318530717958647692f * u1), r0 * std::sin(6.28318530717958647692f * u1),
                  r1 * std::cos(6.28318530717958647692f * u3), r1 * std::sin(6.28318530717958647692f * u3)};
    for (int k = 0; k < 4 && i + k < d; ++k) out[i + k] = (T)z[k];
  }
}

// Tuned CPU transition for the timed baseline only (never used as the checker): the same HMC transition as hmc_step
// (generic_hmc.rs:166-221) without per-call allocation (caller-owned scratch) and, for RosenbrockND, with the gradient in a
// two-pass form the compiler vectorises (t_i = x_{i+1} - x_i^2 first, then g_i from t_i and t_{i-1}) instead of the scalar
// recurrence of the checker.  Same operations per coordinate; summation order of the log density differs (vector lanes).
template <class T>
inline T rosenbrock_logp_grad_fast(const T* __restrict__ x, T* __restrict__ g, T* __restrict__ t, int d) {
  T s = 0;
  for (int i = 0; i < d - 1; ++i) {
    const T ti = x[i + 1] - x[i] * x[i];
    const T u = T(1) - x[i];
    t[i] = ti;
    s += T(100) * ti * ti + u * u;
  }
  g[0] = d > 1 ? T(400) * t[0] * x[0] + T(2) * (T(1) - x[0]) : T(0);
  for (int i = 1; i < d - 1; ++i) g[i] = T(400) * t[i] * x[i] + T(2) * (T(1) - x[i]) - T(200) * t[i - 1];
  if (d > 1) g[d - 1] = T(-200) * t[d - 2];
  return -s;
}

template <class T>
inline int hmc_step_fast(const Target<T>& tgt, T* q, const T* mom, T ln_u, T eps, int L, T* pq, T* pp, T* grad, T* scratch) {
  const int d = tgt.dim;
  const bool rosen = tgt.kind == T_ROSENBROCK_ND;
  auto eval = [&](const T* x) -> T { return rosen ? rosenbrock_logp_grad_fast(x, grad, scratch, d) : tgt.logp_and_grad(x, grad); };
  T ke0 = 0;
  for (int i = 0; i < d; ++i) { pq[i] = q[i]; pp[i] = mom[i]; ke0 += mom[i] * mom[i]; }
  const T logp0 = eval(q);
  const T half = T(0.5) * eps;
  T logp = logp0;
  for (int l = 0; l < L; ++l) {
    for (int i = 0; i < d; ++i) { pp[i] += grad[i] * half; pq[i] += pp[i] * eps; }
    logp = eval(pq);
    for (int i = 0; i < d; ++i) pp[i] += grad[i] * half;
  }
  T ke1 = 0;
  for (int i = 0; i < d; ++i) ke1 += pp[i] * pp[i];
  const T log_accept = (logp - logp0) + (T(0.5) * ke0 - T(0.5) * ke1);
  const int acc = ln_u <= log_accept;
  if (acc) for (int i = 0; i < d; ++i) q[i] = pq[i];
  return acc;
}

template <class T>
double hmc_bench(int kind, int dim, const double* params, size_t np, size_t C, T* q, T eps, int L,
                 size_t n_steps, uint64_t seed, int threads, T* samples) {
  Target<T> t(kind, dim, params, np);
  const size_t d = (size_t)dim;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel
  {
    std::vector<T> mom(d), pq(d), pp(d), grad(d), scratch(d);
#pragma omp for schedule(static)
    for (long long ci = 0; ci < (long long)C; ++ci) {
      size_t c = (size_t)ci;
      for (size_t s = 0; s < n_steps; ++s) {
        philox_normals<T>(seed, c, (uint32_t)s, 0, dim, mom.data());
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t ctr[4] = {(uint32_t)c, (uint32_t)(c >> 32), (uint32_t)s, (1u << 24)};
        uint32_t r[4];
        Philox::block(ctr, key, r);
        T ln_u = (T)std::log(u01_f32(r[0]));
        hmc_step_fast(t, q + c * d, mom.data(), ln_u, eps, L, pq.data(), pp.data(), grad.data(), scratch.data());
        if (samples) std::memcpy(samples + (c * n_steps + s) * d, q + c * d, d * sizeof(T));
      }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

template <class T>
double mh_bench(int kind, int dim, const double* params, size_t np, T prop_std, size_t C, T* x,
                size_t n_steps, uint64_t seed, int threads, double* samples) {
  Target<T> t(kind, dim, params, np);
  const size_t d = (size_t)dim;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel
  {
    std::vector<T> z(d);
#pragma omp for schedule(static)
    for (long long ci = 0; ci < (long long)C; ++ci) {
      size_t c = (size_t)ci;
      for (size_t s = 0; s < n_steps; ++s) {
        philox_normals<T>(seed, c, (uint32_t)s, 0, dim, z.data());
        uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        uint32_t ctr[4] = {(uint32_t)c, (uint32_t)(c >> 32), (uint32_t)s, (1u << 24)};
        uint32_t r[4];
        Philox::block(ctr, key, r);
        T ln_u = (T)std::log(u01_f64(r[0], r[1]));
        mh_step(t, prop_std, x + c * d, z.data(), ln_u);
        if (samples)
          for (size_t k = 0; k < d; ++k) samples[(c * n_steps + s) * d + k] = (double)x[c * d + k];
      }
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // namespace

extern "C" {

double orc_target_logp_grad_f64(int kind, int dim, const double* params, size_t np, const double* x, double* g) { return target_logp_grad<double>(kind, dim, params, np, x, g); }
float orc_target_logp_grad_f32(int kind, int dim, const double* params, size_t np, const float* x, float* g) { return target_logp_grad<float>(kind, dim, params, np, x, g); }
double orc_target_logp_f64(int kind, int dim, const double* params, size_t np, const double* x) { return target_logp<double>(kind, dim, params, np, x); }
float orc_target_logp_f32(int kind, int dim, const double* params, size_t np, const float* x) { return target_logp<float>(kind, dim, params, np, x); }

double orc_iso_proposal_logp_f64(const double* from, const double* to, int d, double s) { return iso_proposal_logp<double>(from, to, d, s); }

void orc_hmc_run_f64(int kind, int dim, const double* params, size_t np, size_t C, double* q, double eps, int L, size_t n_steps, const double* momenta, const double* ln_u, double* samples, uint8_t* accepted, double* log_accept, double* prop_q, double* prop_p, double* logp_cur, double* logp_prop) {
  hmc_run<double>(kind, dim, params, np, C, q, eps, L, n_steps, momenta, ln_u, samples, accepted, log_accept, prop_q, prop_p, logp_cur, logp_prop);
}
void orc_hmc_run_f32(int kind, int dim, const double* params, size_t np, size_t C, float* q, float eps, int L, size_t n_steps, const float* momenta, const float* ln_u, float* samples, uint8_t* accepted, float* log_accept, float* prop_q, float* prop_p, float* logp_cur, float* logp_prop) {
  hmc_run<float>(kind, dim, params, np, C, q, eps, L, n_steps, momenta, ln_u, samples, accepted, log_accept, prop_q, prop_p, logp_cur, logp_prop);
}

void orc_mh_run_f64(int kind, int dim, const double* params, size_t np, double prop_std, size_t C, double* x, size_t n_steps, const double* normals, const double* ln_u, double* samples, uint8_t* accepted, double* log_ratio) {
  mh_run<double>(kind, dim, params, np, prop_std, C, x, n_steps, normals, ln_u, samples, accepted, log_ratio);
}
void orc_mh_run_f32(int kind, int dim, const double* params, size_t np, float prop_std, size_t C, float* x, size_t n_steps, const float* normals, const float* ln_u, double* samples, uint8_t* accepted, float* log_ratio) {
  mh_run<float>(kind, dim, params, np, prop_std, C, x, n_steps, normals, ln_u, samples, accepted, log_ratio);
}

void orc_nuts_build_tree_f64(int kind, int dim, const double* params, size_t np, const double* q, const double* p, const double* g, double logu, int v, int j, double eps, double joint_0, const double* unif, size_t n_unif, double* out_vecs, double* out_scalars, long long* out_ints) {
  build_tree_c<double>(kind, dim, params, np, q, p, g, logu, v, j, eps, joint_0, unif, n_unif, out_vecs, out_scalars, out_ints);
}
void orc_nuts_build_tree_f32(int kind, int dim, const double* params, size_t np, const float* q, const float* p, const float* g, float logu, int v, int j, float eps, float joint_0, const double* unif, size_t n_unif, float* out_vecs, float* out_scalars, long long* out_ints) {
  build_tree_c<float>(kind, dim, params, np, q, p, g, logu, v, j, eps, joint_0, unif, n_unif, out_vecs, out_scalars, out_ints);
}
double orc_nuts_find_reasonable_epsilon_f64(int kind, int dim, const double* params, size_t np, const double* q, const double* p) {
  Target<double> t(kind, dim, params, np);
  return nuts_find_reasonable_epsilon<double>(t, q, p);
}
float orc_nuts_find_reasonable_epsilon_f32(int kind, int dim, const double* params, size_t np, const float* q, const float* p) {
  Target<float> t(kind, dim, params, np);
  return nuts_find_reasonable_epsilon<float>(t, q, p);
}
void orc_nuts_run_f64(int kind, int dim, const double* params, size_t np, size_t C, double* q, double target_accept, int max_depth, double eps_init, size_t n_collect, size_t n_discard, const double* normals, size_t n_norm, const double* exp1, size_t n_exp, const double* unif, size_t n_unif, double* samples, double* eps_final, long long* leapfrogs, long long* used, int* exhausted) {
  nuts_run<double>(kind, dim, params, np, C, q, target_accept, max_depth, eps_init, n_collect, n_discard, normals, n_norm, exp1, n_exp, unif, n_unif, samples, eps_final, leapfrogs, used, exhausted);
}
void orc_nuts_run_f32(int kind, int dim, const double* params, size_t np, size_t C, float* q, float target_accept, int max_depth, float eps_init, size_t n_collect, size_t n_discard, const double* normals, size_t n_norm, const double* exp1, size_t n_exp, const double* unif, size_t n_unif, float* samples, float* eps_final, long long* leapfrogs, long long* used, int* exhausted) {
  nuts_run<float>(kind, dim, params, np, C, q, target_accept, max_depth, eps_init, n_collect, n_discard, normals, n_norm, exp1, n_exp, unif, n_unif, samples, eps_final, leapfrogs, used, exhausted);
}
// with diagonal mass-matrix adaptation (GenericNUTS::new_with_mass_matrix, generic_nuts.rs:379-398)
void orc_nuts_run_mass_f64(int kind, int dim, const double* params, size_t np, size_t C, double* q, double target_accept, int max_depth, double eps_init, size_t n_collect, size_t n_discard, const double* normals, size_t n_norm, const double* exp1, size_t n_exp, const double* unif, size_t n_unif, double* samples, double* eps_final, long long* leapfrogs, long long* used, int* exhausted, const double* mass_cfg, double* mass_inv_out, long long* mass_updates_out) {
  nuts_run<double>(kind, dim, params, np, C, q, target_accept, max_depth, eps_init, n_collect, n_discard, normals, n_norm, exp1, n_exp, unif, n_unif, samples, eps_final, leapfrogs, used, exhausted, mass_cfg, mass_inv_out, 1, mass_updates_out);
}
void orc_nuts_run_mass_f32(int kind, int dim, const double* params, size_t np, size_t C, float* q, float target_accept, int max_depth, float eps_init, size_t n_collect, size_t n_discard, const double* normals, size_t n_norm, const double* exp1, size_t n_exp, const double* unif, size_t n_unif, float* samples, float* eps_final, long long* leapfrogs, long long* used, int* exhausted, const double* mass_cfg, float* mass_inv_out, long long* mass_updates_out) {
  nuts_run<float>(kind, dim, params, np, C, q, target_accept, max_depth, eps_init, n_collect, n_discard, normals, n_norm, exp1, n_exp, unif, n_unif, samples, eps_final, leapfrogs, used, exhausted, mass_cfg, mass_inv_out, 1, mass_updates_out);
}

// ---- mass matrix (MassMatrix::diagonal_from_var / kinetic / inv_mul, generic_nuts.rs:196-206, 228-281) ----
// Reference KAT generic_nuts.rs:1427-1440: var = [4, 9], p = [2, 3] -> kinetic = 1.0, inv_mul = [0.5, 1/3].
// Reference KAT generic_nuts.rs:1442-1457 (dense_mass_matrix_inverse_matches_identity_action): cov = [[2, .3], [.3, 1]], p = [.7, -1.1]:
// MassMatrix::dense_from_cov -> inv_mul; returns 1 when the factorisation succeeded.  Also hands back inv and chol.
int orc_dense_mass_inv_mul_f64(const double* cov, int d, double jitter, const double* p, double* inv_mul_out, double* inv_out, double* chol_out, double* ke_out) {
  DiagMass<double> m;
  if (!DiagMass<double>::dense_from_cov(std::vector<double>(cov, cov + (size_t)d * d), d, jitter, &m)) return 0;
  m.inv_mul(p, inv_mul_out, d);
  if (inv_out) std::copy(m.inv.begin(), m.inv.end(), inv_out);
  if (chol_out) std::copy(m.chol.begin(), m.chol.end(), chol_out);
  if (ke_out) *ke_out = nuts_kinetic<double>(p, d, &m);
  return 1;
}
double orc_diag_mass_kinetic_inv_mul_f64(const double* var, int d, double jitter, const double* p, double* inv_mul_out) {
  DiagMass<double> m = DiagMass<double>::from_var(std::vector<double>(var, var + d), jitter);
  for (int i = 0; i < d; ++i) inv_mul_out[i] = m.inv[i] * p[i];   // inv_mul, :265-281
  return nuts_kinetic<double>(p, d, &m);
}

// ---- integer-state MH (tests/metrohast_poisson_test.rs) ----
// x [C,d] int32 in/out; steps [n,C,d] int8 (+1 / -1); ln_u [n,C]; samples f64 [C,n,d]; accepted [n,C]; log_ratio [n,C]
void orc_mh_int_run(int kind, int dim, const double* params, size_t C, int* x, size_t n_steps, const signed char* steps,
                    const double* ln_u, double* samples, uint8_t* accepted, double* log_ratio) {
  IntTarget t;
  t.kind = kind; t.dim = dim;
  if (kind == IT_POISSON) t.lambda = params[0]; else { t.n = (int)params[0]; t.p = params[1]; }
  const size_t d = (size_t)dim;
#pragma omp parallel for schedule(static)
  for (long long ci = 0; ci < (long long)C; ++ci) {
    const size_t c = (size_t)ci;
    for (size_t s = 0; s < n_steps; ++s) {
      MhIntStepInfo r = mh_int_step(t, x + c * d, steps + (s * C + c) * d, ln_u[s * C + c]);
      if (samples) for (size_t k = 0; k < d; ++k) samples[(c * n_steps + s) * d + k] = (double)x[c * d + k];
      if (accepted) accepted[s * C + c] = (uint8_t)r.accepted;
      if (log_ratio) log_ratio[s * C + c] = r.log_accept_ratio;
    }
  }
}
double orc_int_target_logp(int kind, int dim, const double* params, const int* k) {
  IntTarget t;
  t.kind = kind; t.dim = dim;
  if (kind == IT_POISSON) t.lambda = params[0]; else { t.n = (int)params[0]; t.p = params[1]; }
  return t.logp(k);
}

// ---- Gibbs sweeps (gibbs.rs:89-105) ----
// x [C,d] f64 in/out; normals, uniforms [n,C,d]; samples f64 [C,n,d]
void orc_gibbs_run(int kind, int dim, const double* params, size_t C, double* x, size_t n_steps, const double* normals,
                   const double* uniforms, double* samples) {
  GibbsConditional g;
  g.kind = kind;
  if (kind == GC_CONSTANT) g.c = params[0];
  else { g.mu0 = params[0]; g.sigma0 = params[1]; g.mu1 = params[2]; g.sigma1 = params[3]; g.pi0 = params[4]; }
  const size_t d = (size_t)dim;
#pragma omp parallel for schedule(static)
  for (long long ci = 0; ci < (long long)C; ++ci) {
    const size_t c = (size_t)ci;
    for (size_t s = 0; s < n_steps; ++s) {
      gibbs_step(g, x + c * d, dim, normals + (s * C + c) * d, uniforms + (s * C + c) * d);
      if (samples) for (size_t k = 0; k < d; ++k) samples[(c * n_steps + s) * d + k] = x[c * d + k];
    }
  }
}

// ---- stats ----
void orc_split_rhat_mean_ess(const float* sample, size_t c, size_t n, size_t p, float* rhat, float* ess_out) { split_rhat_mean_ess(sample, c, n, p, rhat, ess_out); }
void orc_autocov_bf(const float* x, size_t n, size_t d, float* out) { autocov_bf(x, n, d, out); }
void orc_autocov_fft(const float* x, size_t n, size_t d, float* out) { autocov_fft(x, n, d, out); }
void orc_basic_stats(const float* data, size_t n, float* out5) {
  BasicStats b = basic_stats(std::vector<float>(data, data + n));
  out5[0] = b.min; out5[1] = b.median; out5[2] = b.max; out5[3] = b.mean; out5[4] = b.std;
}
// feeds `steps` states [steps, c, p] to a MultiChainTracker and returns rhat[p], p_accept
void orc_tracker_rhat(const float* states, size_t steps, size_t c, size_t p, float* rhat, float* p_accept) {
  MultiChainTracker tr(c, p);
  for (size_t s = 0; s < steps; ++s) tr.step(states + s * c * p);
  tr.rhat(rhat);
  if (p_accept) *p_accept = tr.p_accept;
}

// ---- Philox (host restatement of the product's RNG contract) ----
void orc_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { Philox::block(ctr, key, out); }

// ---- timed CPU baseline ("port") ----
double orc_hmc_bench_f32(int kind, int dim, const double* params, size_t np, size_t C, float* q, float eps, int L, size_t n_steps, uint64_t seed, int threads, float* samples) { return hmc_bench<float>(kind, dim, params, np, C, q, eps, L, n_steps, seed, threads, samples); }
double orc_hmc_bench_f64(int kind, int dim, const double* params, size_t np, size_t C, double* q, double eps, int L, size_t n_steps, uint64_t seed, int threads, double* samples) { return hmc_bench<double>(kind, dim, params, np, C, q, eps, L, n_steps, seed, threads, samples); }
double orc_mh_bench_f64(int kind, int dim, const double* params, size_t np, double prop_std, size_t C, double* x, size_t n_steps, uint64_t seed, int threads, double* samples) { return mh_bench<double>(kind, dim, params, np, prop_std, C, x, n_steps, seed, threads, samples); }
void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
