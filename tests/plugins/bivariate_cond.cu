// Test plugin: Gibbs conditionals of a standard bivariate normal with correlation rho = params[0]:
//   x_i | x_j ~ N(rho x_j, 1 - rho^2)
#include "gmcmc_custom_conditional.cuh"

struct BivariateNormal {
  static constexpr int dim = 2;
  template <class RNG>
  __device__ static double sample(int i, const double (&given)[dim], const double* params, RNG& rng) {
    const double rho = params[0];
    return rho * given[1 - i] + sqrt(1.0 - rho * rho) * rng.normal();
  }
};
GMCMC_REGISTER_CONDITIONAL(BivariateNormal)
