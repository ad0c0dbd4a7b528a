"""Fit the BNN's variational posterior on the GPU (vihmc.vi.train_bbb, the reference configuration: 10 000 epochs x 10 draws) and
select the 40 most sensitive weights (vihmc.sensitivity): the fitted start of the ESS leg of bench.py and of the long-run
posterior-parity test.  Writes gpurun_out/bnn_vi_fit.npz (copied to tests/golden/ -- the CPU side cannot run the trainer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import numpy as np
from vihmc import sensitivity, synth, vi
from vihmc.spec import LogProbSpec, MLPArch

x, y, xv, yv = synth.bnn_data()
arch = MLPArch(in_dim=1, widths=(10, 10), out_dim=1, act="tanh", last_bias=True)
mk = lambda a_, b_: LogProbSpec(arch=arch, x=a_, y=b_, loss="NLL", tau_out=0.05 ** 2, prior_sigma_scalar=1.0)
fit = vi.train_bbb(mk(x, y), mk(xv, yv), epochs=10_000, num_ens=10, lr_start=1e-2, lr_patience=5000, seed=0)
scores = np.asarray(sensitivity.eval_std_dydw((xv, None), arch, fit.best_mu, fit.best_sigma))
ind = np.sort(np.argsort(-scores)[:40]).astype(np.int64)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "bnn_vi_fit.npz"), mu=fit.best_mu.numpy(), sigma=fit.best_sigma.numpy(), ind=ind, scores=scores.astype(np.float32))
print("ok", fit.best_mu.shape, ind[:8], float(fit.history[-1, 0]))
