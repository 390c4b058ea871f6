import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, time
from magi_v2_b200 import MAGI_v2
g = np.load("tests/golden/seir_datasets.npz")
X = g["X_obs"][0][:, 1:].copy(); X[X < 0] = 0
m = MAGI_v2(3, g["ts_obs"], X, 80, "seir3")
t = time.time(); m.initial_fit(1); print("fit s", time.time() - t)
print("phi1", m.phi1s, "phi2", m.phi2s, "sig2", m.sigma_sqs_init, "theta_init", m.thetas_init)
for L, nb, nr in ((32, 600, 600), (128, 1000, 1000), (256, 1000, 1000)):
    t = time.time()
    r = m.predict(num_results=nr, num_burnin_steps=nb, n_chains=8, n_leapfrog=L, seed=3)
    th = r["thetas_samps"]
    kr = r["kernel_results"]
    print(f"L={L} burn={nb} res={nr}: {time.time()-t:.1f}s eps={kr['step_size'][:3]} acc={kr['accept_prob'].mean():.2f}")
    print("   mean first 100:", th[:, :100].mean(axis=(0, 1)), " last 100:", th[:, -100:].mean(axis=(0, 1)), " all:", th.mean(axis=(0, 1)))
    print("   sigma2 mean", r["sigma_sqs_samps"].mean(axis=(0, 1)))
