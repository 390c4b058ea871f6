import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, time
from magi_v2_b200 import MAGI_v2
g = np.load("tests/golden/seir_datasets.npz")
X = g["X_obs"][0][:, 1:].copy(); X[X < 0] = 0
Xt = g["X_true"][0][:, 1:]
sig2_true = (0.05 * (Xt.max(0) - Xt.min(0))) ** 2
print("true noise var", sig2_true)
for name, hp in (("survey-phi + true sigma", dict(phi1s=[0.0085, 0.034, 0.024], phi2s=[0.375, 0.23, 0.109], sigma_sqs=sig2_true)),
                 ("fitted", None)):
    m = MAGI_v2(3, g["ts_obs"], X, 80, "seir3")
    m.initial_fit(1, hparams=hp)
    print("==", name, "phi1", m.phi1s, "phi2", m.phi2s, "sig2", m.sigma_sqs_init, "theta_init", m.thetas_init)
    for L, nb, nr in ((64, 1000, 1000),):
        t = time.time()
        r = m.predict(num_results=nr, num_burnin_steps=nb, n_chains=8, n_leapfrog=L, seed=3)
        th = r["thetas_samps"]; kr = r["kernel_results"]
        print(f"L={L}: {time.time()-t:.1f}s eps={kr['step_size'][:2]} acc={kr['accept_prob'].mean():.2f}")
        print("   mean first 100:", th[:, :100].mean(axis=(0, 1)), " last 100:", th[:, -100:].mean(axis=(0, 1)))
        print("   per-chain means beta:", th[:, :, 0].mean(axis=1))
        print("   sigma2 mean", r["sigma_sqs_samps"].mean(axis=(0, 1)))
