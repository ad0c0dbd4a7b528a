"""How long do the BNN VI-HMC chains of BASELINE configs[1] need to reach stationarity, and what is ESS/s there?
VI fit (vihmc.vi.train_bbb) -> sensitivity scores -> the d most sensitive coordinates -> 1024 chains started from the fitted
variational posterior -> segments of `--segment` HMC iterations, each continuing from the last state of the previous one;
per segment: acceptance, rank-normalised split-R-hat (max / median over coordinates), bulk-ESS (min / median), wall time."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from vihmc import diagnostics, engine, samplers, sensitivity, synth, vi  # noqa: E402
from vihmc.spec import LogProbSpec, MLPArch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=10_000)
    ap.add_argument("--chains", type=int, default=1024)
    ap.add_argument("--d", type=int, default=40)
    ap.add_argument("--segment", type=int, default=500)
    ap.add_argument("--segments", type=int, default=8)
    ap.add_argument("--L", type=int, default=196)
    ap.add_argument("--eps", type=float, default=5e-4)
    ap.add_argument("--thin", type=int, default=1)
    a = ap.parse_args()
    x, y, xv, yv = synth.bnn_data()
    arch = MLPArch(in_dim=1, widths=(10, 10), out_dim=1, act="tanh", last_bias=True)
    mk = lambda a_, b_: LogProbSpec(arch=arch, x=a_, y=b_, loss="NLL", tau_out=0.05 ** 2, prior_sigma_scalar=1.0)
    t0 = time.perf_counter()
    fit = vi.train_bbb(mk(x, y), mk(xv, yv), epochs=a.epochs, num_ens=10, lr_start=1e-2, lr_patience=5000, seed=0)
    scores = sensitivity.eval_std_dydw((xv, None), arch, fit.best_mu, fit.best_sigma)
    ind = np.sort(np.argsort(-np.asarray(scores))[:a.d]).astype(np.int64)
    mu, sigma = fit.best_mu, fit.best_sigma
    numels = arch.tensor_numels()
    spec = samplers.define_model_log_prob_bnn(arch, "NLL", x, y, numels, None, [torch.tensor(1.0) for _ in numels], 0.0025,
                                              params_mu=mu, params_std=sigma, grad_ind=ind)
    print(json.dumps({"vi_fit_and_selection_s": time.perf_counter() - t0, "d": int(len(ind)), "sigma_sel_median": float(sigma[ind].median())}))
    g = torch.Generator().manual_seed(0)
    q = (mu[ind][None] + sigma[ind][None] * torch.randn(a.chains, len(ind), generator=g)).cuda()
    prep = engine.prepare(spec)
    # thinned long run: chunks of `thin` iterations (burn = thin - 2 keeps row 0 = the chunk's start and row 1 = its end state)
    kept, accs, lps = [], [], []
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    n_chunks = a.segments * a.segment
    for k in range(n_chunks):
        res = engine.run_sampler([prep], q, a.thin, a.L, a.eps, burn=a.thin - 2, seed=1000 + k, diagnostics=True, to_host=False,
                                 hamiltorch_fallback_rule=False)
        q = res.samples[-1]
        kept.append(q)
        lps.append(res.logp[-1])
        accs.append(res.accepted.float().mean())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t1
    draws, lp = torch.stack(kept), torch.stack(lps)
    print(json.dumps({"iterations": n_chunks * a.thin, "thin": a.thin, "seconds": dt, "accept": float(torch.stack(accs).mean()),
                      "evals_per_s": a.chains * n_chunks * a.thin * (a.L + 1) / dt}))
    for s in range(a.segments):
        w, wl = draws[s * a.segment:(s + 1) * a.segment], lp[s * a.segment:(s + 1) * a.segment]
        summ = diagnostics.summarize(w, logp=wl)
        print(json.dumps({"window": s, "thinned_draws": a.segment, "rhat_max": summ["rhat_max"], "rhat_median": summ["rhat_median"],
                          "ess_min": summ["ess_bulk_min"], "ess_median": summ["ess_bulk_median"], "ess_logp": summ["ess_bulk_logp"],
                          "rhat_logp": summ["rhat_logp"], "logp_mean": float(wl.mean()),
                          "ess_min_per_s": summ["ess_bulk_min"] / (dt / a.segments)}), flush=True)


if __name__ == "__main__":
    main()
