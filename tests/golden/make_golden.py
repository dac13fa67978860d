"""Generates tests/golden/*.json.

Two kinds of fixtures:
  reference_kats.json  — the RNG-free known-answer vectors the REFERENCE's own tests hold for this path, copied as
                         numbers with their source (file:line into /root/reference/src).  The reference is Rust and
                         cannot be run in this image, so these are transcribed, not regenerated.
  oracle_cfg1.json     — outputs of the CPU oracle (oracle/gmcmc_oracle.hpp) on BASELINE config 1 with the fixed
                         inputs below; the GPU tests compare the CUDA path against these committed numbers in
                         addition to the live oracle (so a drifting oracle is caught too).
Run:  python tests/golden/make_golden.py     (needs gcc for the oracle; no GPU)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

REFERENCE_KATS = {
    "nuts_build_tree": {
        "source": "nuts.rs:521-586 (test_build_tree): DiffableGaussian2D([0,1],[[4,2],[2,3]]), q=[0,1], p=[2,3], g=[4,5], "
                  "logu=-2, v=-1, j=3, eps=0.01, joint_0=0.1; tolerance rel 1e-5 / abs 1e-6",
        "q_minus": [-0.1584001, 0.76208336], "p_minus": [1.9800036, 2.9718253], "g_minus": [-7.91236e-5, 7.9358295e-2],
        "q_plus": [-0.0198, 0.97025], "p_plus": [1.98, 2.9749503], "g_plus": [-1.250e-05, 9.925e-03],
        "q_prime": [-0.0198, 0.97025], "g_prime": [-1.250e-05, 9.925e-03],
        "n_prime": 0, "s_prime": True, "n_alpha_prime": 8, "logp_prime": -2.8777454, "alpha_prime": 0.0006866617},
    "find_reasonable_epsilon": {"source": "nuts.rs:508-519: N(0,I), q=[0,1], p=[1,0]", "epsilon": 2.0},
    "nuts_run_1_0": {"source": "nuts.rs:588-601 (test_chain_1): run(1,0) returns the initial point", "sample": [0.0, 1.0]},
    "tracker_rhat": {"source": "stats.rs:734-783", "values_a": [1.4142135, 1.0801234, 0.8944273, 0.8660254],
                     "values_b": [0.70710677, 0.74535599, 1.0, 1.5]},
    "autocov": {"source": "stats.rs:808-839 (brute force and FFT, tol 1e-6)",
                "values": [[1.25, 6.516875], [0.3125, -3.7889063], [-0.375, 1.4721875], [-0.5625, -0.94171875]]},
    "gaussian2d_logp": {"source": "distributions.rs:820-839: identity cov, x=(0.5,-0.5)", "value": -2.0878770664093453},
    "iso_gauss_density": {"source": "distributions.rs:580-614", "values": [0.24197072451914337, 3.864661987252467e-7, 0.001080393185560214]},
}

CFG1_START = [[0.30471708, -1.03998411, 0.7504512], [0.94056472, -1.95103519, -1.30217951],
              [0.1278404, -0.31624259, -0.01680116], [-0.85304393, 0.87939797, 0.77779194]]


def main():
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(REFERENCE_KATS, f, indent=1)
    out = {"config": "examples/rosenbrock3d_hmc: RosenbrockND d=3, 4 chains, eps=0.01, L=10, 50 discard + 400 collect; "
                     "momenta = default_rng(42).standard_normal((450,4,3)), ln u = log(default_rng(42)...random((450,4)))",
           "start": CFG1_START}
    for name, dt in (("f32", np.float32), ("f64", np.float64)):
        rng = np.random.default_rng(42)
        mom = rng.standard_normal((450, 4, 3)).astype(dt)
        ln_u = np.log(rng.random((450, 4))).astype(dt)
        r = O.hmc_run(O.ROSENBROCK_ND, [], np.asarray(CFG1_START, dt), 0.01, 10, mom, ln_u)
        out[name] = {"final_positions": r["q"].astype(np.float64).tolist(),
                     "accepted_per_chain": r["accepted"].sum(0).astype(int).tolist(),
                     "sample_399": r["samples"][:, -1, :].astype(np.float64).tolist(),
                     "sample_mean": r["samples"][:, 50:, :].astype(np.float64).mean(axis=(0, 1)).tolist()}
    with open(os.path.join(HERE, "oracle_cfg1.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
