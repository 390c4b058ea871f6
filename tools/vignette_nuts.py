"""The reference's vignette (vignette.ipynb: SEIR, E/I/R observed, discretization 1, band 80, 1000 + 1000 NUTS
iterations) through MAGI_v2.predict(sampler="nuts"); prints the posterior means next to the vignette's printed ones
(vignette.ipynb:281-283: 5.831, 0.565, 1.77) and the truth (6.0, 0.6, 1.8).
    python tools/vignette_nuts.py [n_chains] [num_results] [num_burnin]"""
import sys
import time

import numpy as np

from magi_v2_b200 import MAGI_v2
from tests.helpers import load_golden

R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nres = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
nburn = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
bt = float(sys.argv[4]) if len(sys.argv) > 4 else None      # fixed beta_temp (default: the reference's schedule)
depth = int(sys.argv[5]) if len(sys.argv) > 5 else 10
g = load_golden("seir_datasets.npz")
X = g["X_obs"][0][:, 1:].copy()
X[X < 0.0] = 0.0
t0 = time.time()
model = MAGI_v2(D_thetas=3, ts_obs=g["ts_obs"], X_obs=X, bandsize=80, f_vec="seir3")
model.initial_fit(discretization=1, verbose=False)
t1 = time.time()
print(f"initial_fit {t1 - t0:.1f} s; phi1 {model.phi1s}, phi2 {model.phi2s}, sigma_sq {model.sigma_sqs_init}, theta_init {model.thetas_init}")
res = model.predict(num_results=nres, num_burnin_steps=nburn, sampler="nuts", n_chains=R, seed=0, beta_temp=bt,
                    max_tree_depth=depth)
t2 = time.time()
th = res["thetas_samps"].reshape(R, nres, 3)
kr = res["kernel_results"]
lf = np.asarray(kr["leapfrogs_taken"]).reshape(R, nres)
print(f"predict {t2 - t1:.1f} s; step sizes {kr['step_size']}; mean leapfrogs/transition {lf.mean():.1f} (max {lf.max()}); "
      f"accept {np.asarray(kr['accept_prob']).mean():.3f}; divergences {int(np.asarray(kr['has_divergence']).sum())}")
for r in range(R):
    print(f"chain {r}: theta mean {th[r].mean(0)}  sd {th[r].std(0)}")
print("all chains: mean", th.mean((0, 1)), " sd", th.reshape(-1, 3).std(0), " between-chain sd of means", th.mean(1).std(0))
print("vignette.ipynb:281-283 means: [5.831 0.565 1.77]; truth [6.0 0.6 1.8]")
print("sigma_sq means", res["sigma_sqs_samps"].reshape(-1, 3).mean(0))
