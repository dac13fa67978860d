import sys, numpy as np
sys.path.insert(0, '/root/repo')
import general_mcmc_b200 as gm
DIM=100; K=4; chains=65536
mu = np.stack([(k - 1.5) * (2.0 / np.sqrt(DIM)) * np.ones(DIM) for k in range(K)])
tgt = gm.GaussianMixture(np.full(K, 1.0 / K), mu, 1.0)
q0 = np.random.default_rng(300).standard_normal((chains, DIM)).astype(np.float32)
s = gm.NUTS(tgt, q0, 0.8, seed=42, max_depth=10)
s.run_device(1, 100)
st0 = s.state()
s.run_device(21, 0)
st1 = s.state()
lf = st1["leapfrogs"] - st0["leapfrogs"]
print("eps percentiles", np.percentile(st1["eps"], [0, 0.1, 1, 50, 99, 100]))
print("leapfrogs per chain over 20 transitions: percentiles", np.percentile(lf, [0, 50, 90, 99, 99.9, 100]), "mean", lf.mean())
print("sum", lf.sum(), "max*chains", lf.max() * chains, "balance", lf.sum() / (lf.max() * chains))
