// Example custom target (test fixture): a 3-D "banana" — x0 ~ N(0, s^2), x1 | x0 ~ N(b x0^2, 1), x2 ~ N(1, 0.25).
// params = [s, b].  Compiled ahead of time into tests/plugins/banana.so (general_mcmc_b200.build_custom_target).
#include "gmcmc_custom_target.cuh"

struct Banana {
  static constexpr int dim = 3;
  template <class T>
  __device__ static T logp_grad(const T (&x)[3], T (&g)[3], const T* params) {
    const T s = params[0], b = params[1];
    const T r = x[1] - b * x[0] * x[0];
    const T u = x[2] - T(1);
    g[0] = -x[0] / (s * s) + T(2) * b * x[0] * r;
    g[1] = -r;
    g[2] = -T(4) * u;
    return -T(0.5) * x[0] * x[0] / (s * s) - T(0.5) * r * r - T(2) * u * u;
  }
};
GMCMC_REGISTER_CUSTOM_TARGET(Banana)
