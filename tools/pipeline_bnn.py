"""The reference's whole BNN workflow on the engine, end to end on one GPU:
   1. Bayes-by-Backprop fit            Neural_network/VI/main_regression_VI.py      -> vihmc.vi.train_bbb
   2. sensitivity scores + selection   Neural_network/VI/sensitivity.py             -> vihmc.sensitivity.eval_std_dydw / select_indices
   3. artefact files                   means_flattened / stds_flattened / gradient_indices
   4. VI-HMC over the selected subset  Neural_network/VI_HMC/main_VI_HMC.py         -> vihmc.samplers.sample (1024 chains)
   5. posterior prediction             main_VI_HMC.py:384-429 validate()            -> vihmc.samplers.predict_model
and the diagnostics (split-R-hat, bulk-ESS) of the chains -- with a FITTED variational posterior the chains start in the typical set,
unlike the random synthetic artefacts SURVEY 8(d) prescribes for the bench workload."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from vihmc import artifacts, diagnostics, samplers, sensitivity, synth, vi  # noqa: E402
from vihmc.spec import LogProbSpec, MLPArch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=10_000)
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--samples", type=int, default=300)
    ap.add_argument("--L", type=int, default=196)
    ap.add_argument("--eps", type=float, default=5e-4)
    ap.add_argument("--threshold", type=float, default=0.90)
    a = ap.parse_args()
    x, y, xv, yv = synth.bnn_data()
    arch = MLPArch(in_dim=1, widths=(10, 10), out_dim=1, act="tanh", last_bias=True)
    noise_var = 0.05 ** 2                                        # Neural_network/VI/config.py: noise = 5e-2 (std)
    mk = lambda a_, b_: LogProbSpec(arch=arch, x=a_, y=b_, loss="NLL", tau_out=noise_var, prior_sigma_scalar=1.0)
    out = {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit = vi.train_bbb(mk(x, y), mk(xv, yv), epochs=a.epochs, num_ens=10, lr_start=1e-2, lr_patience=5000, seed=0)
    t1 = time.perf_counter()
    scores = sensitivity.eval_std_dydw((xv, None), arch, fit.best_mu, fit.best_sigma)
    ind = sensitivity.select_indices(scores, a.threshold)
    t2 = time.perf_counter()
    with tempfile.TemporaryDirectory() as tmp:
        artifacts.save_vi_artifacts(tmp, "pipe", fit.best_mu, fit.best_sigma, ind, scores)
        mu, sigma, ind = artifacts.load_vi_artifacts(tmp, "pipe")
    d = len(ind)
    numels = arch.tensor_numels()
    spec = samplers.define_model_log_prob_bnn(arch, "NLL", x, y, numels, None, [torch.tensor(1.0) for _ in numels], 0.0025,
                                              params_mu=mu, params_std=sigma, grad_ind=ind)
    g = torch.Generator().manual_seed(0)
    q0 = mu[ind][None] + sigma[ind][None] * torch.randn(a.chains, d, generator=g)
    t3 = time.perf_counter()
    res = samplers.sample(spec, q0, num_samples=a.samples, num_steps_per_sample=a.L, step_size=a.eps, num_chains=a.chains, seed=1,
                          return_result=True)
    t4 = time.perf_counter()
    burn = a.samples // 5
    thin = res.samples[burn::10, :64].reshape(-1, d)
    pred, logp = samplers.predict_model(spec, thin, x=xv, y=yv)
    t5 = time.perf_counter()
    summ = diagnostics.summarize(res.samples[burn:].cuda(), logp=res.logp[burn:].cuda())   # rank statistics as torch ops on the GPU
    torch.cuda.synchronize()
    t6 = time.perf_counter()
    pred_vi, _ = samplers.predict_model(spec, mu[ind][None], x=xv, y=yv)
    pm = pred.mean(0).squeeze(-1)
    out = {"vi_fit_s": t1 - t0, "vi_train_loss_first_last": [float(fit.history[0, 0]), float(fit.history[-1, 0])],
           "sensitivity_s": t2 - t1, "selected_d": d, "of_D": int(arch.num_params), "captured_variance": a.threshold,
           "hmc_s": t4 - t3, "chains": a.chains, "samples": a.samples, "L": a.L, "step_size": a.eps,
           "grad_evals_per_s": a.chains * a.samples * (a.L + 1) / (t4 - t3), "acceptance_rate": res.acceptance_rate,
           "rhat_max": float(summ["rhat_max"]), "ess_bulk_min": float(summ["ess_bulk_min"]), "ess_bulk_median": float(summ["ess_bulk_median"]),
           "ess_min_per_s": float(summ["ess_bulk_min"]) / (t4 - t3), "post_burn_draws_per_chain": int(summ["draws"]),
           "predict_s": t5 - t4, "predicted_samples": int(thin.shape[0]), "diagnostics_s": t6 - t5,
           "posterior_mean_mse_on_validation": float(((pm - yv.squeeze(-1)) ** 2).mean()),
           "vi_mean_mse_on_validation": float(((pred_vi[0].squeeze(-1) - yv.squeeze(-1)) ** 2).mean())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
