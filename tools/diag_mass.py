import sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import general_mcmc_b200 as gm
import oracle_lib as oracle
ctx = gm.default_context()
Cn, d, n_collect = 96, 5, 6
scales = np.array([0.3, 1.0, 3.0, 0.7, 2.0])
tgt = gm.DenseGaussian(np.zeros(d), cov=np.diag(scales ** 2))
def streams(Cn, d, steps, seed, n_unif):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((Cn, d * (steps + 2))), rng.exponential(size=(Cn, steps + 2)), rng.random((Cn, n_unif))
for dtype in (np.float64, np.float32):
  for n_discard in (1, 2, 3, 6, 13, 14, 40):
    for cfg in (None, (3, 2, 10, 0.05, 1e-6)):
        rng = np.random.default_rng(21)
        q0 = (rng.standard_normal((Cn, d)) * scales).astype(dtype)
        normals, exp1, unif = streams(Cn, d, n_collect + n_discard + 4, 23, 20000)
        ref = oracle.nuts_run(tgt.kind, tgt.params(), q0, 0.8, 8, -1.0, n_collect, n_discard, normals, exp1, unif, mass_cfg=cfg)
        mm = gm.NUTSMassMatrixConfig("diagonal", *cfg) if cfg else None
        s = gm.NUTS(tgt, q0, 0.8, seed=1, ctx=ctx, max_depth=8, mass_matrix=mm).set_math_mode(True)
        s.inject_streams(normals, exp1, unif)
        out = s.run(n_collect, n_discard)
        st = s.state()
        same = st["leapfrogs"] == ref["leapfrogs"]
        rel = lambda a, b: (np.abs(a - b) / (np.abs(b) + 1e-300)).max() if a.size else 0
        line = "%s n_discard=%d mass=%s same=%.2f eps_rel=%.2e samp_rel=%.2e" % (dtype.__name__, n_discard, bool(cfg), same.mean(), rel(st["eps"][same], ref["eps"][same]), np.abs(out[same] - ref["samples"][same]).max())
        if cfg:
            inv, n = s.mass_matrix()
            line += " n_upd=%d inv_rel=%.2e (per-chain median %.2e)" % (n, rel(inv[same], ref["mass_inv"][same]), np.median((np.abs(inv[same] - ref["mass_inv"][same]) / ref["mass_inv"][same]).max(1)))
        print(line, flush=True)
