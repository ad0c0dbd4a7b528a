#!/usr/bin/env python
"""Time the DeepONet log-posterior gradient (BASELINE.json configs[2]/[3] shape) on one B200.

  python tools/bench_deeponet.py --chains 64 [--n 1000 --nt 101 --nx 101] [--vi] [--reps 3]

Prints one JSON line: chain-grad-evals/s and FP32-equivalent TFLOP/s (11.56 GFLOP per unit at the
shipped shape, SURVEY.md 8(d): fwd 3.863 + bwd 7.695 GFLOP)."""
import argparse, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import numpy as np
import torch
from vihmc import engine, synth
from vihmc.spec import DeepONetArch, LogProbSpec


def flops_per_eval(arch, N, P):
    def stack(dims, rows):
        f = 0
        for li, (o, i) in enumerate(dims):
            f += 2 * rows * o * i * (3 if li > 0 else 2)   # fwd + dW (+ dX except for the first layer)
        return f
    K = arch.output_neurons
    return stack(arch.stack_dims("branch"), N) + stack(arch.stack_dims("trunk"), P) + 3 * 2 * N * P * K


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=32)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--nt", type=int, default=101)
    ap.add_argument("--nx", type=int, default=101)
    ap.add_argument("--vi", action="store_true", help="VI-HMC split: sample a random 10 %% of the weights")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    arch = DeepONetArch()
    rs = np.random.RandomState(0)
    P = a.nt * a.nx
    x1, x2, y, theta = synth.burgers_like(arch, n_train=a.n, n_t=a.nt, n_x=a.nx, seed=0)   # SURVEY 8(d) cfg3 data
    kw = dict(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1)
    if a.vi:
        mu, sigma, ind = synth.deeponet_vi_artifacts(theta, 0.10, seed=1)
        spec = LogProbSpec(frozen=mu, sens_ind=ind, **kw)
        q = mu[ind][None].repeat(a.chains, 1)
    else:
        spec = LogProbSpec(**kw)
        q = theta[None].repeat(a.chains, 1)
    q = (q + 0.01 * torch.from_numpy(rs.randn(*q.shape).astype(np.float32))).cuda()
    prep = engine.prepare(spec)
    engine.logp_grad(prep, q)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.reps + 1)]
    ev[0].record()
    for r in range(a.reps):
        logp, grad = engine.logp_grad(prep, q)
        ev[r + 1].record()
    torch.cuda.synchronize()
    per_rep = [ev[r].elapsed_time(ev[r + 1]) for r in range(a.reps)]
    ms = ev[0].elapsed_time(ev[-1]) / a.reps
    fl = flops_per_eval(arch, a.n, P)
    print(json.dumps({"workload": f"deeponet logp_grad N={a.n} P={P} D={arch.num_params} d={spec.d} chains={a.chains}",
                      "ms_per_eval_batch": ms, "ms_per_rep": [round(t, 2) for t in per_rep], "chain_grad_evals_per_s": a.chains / (ms * 1e-3),
                      "gflop_per_unit": fl / 1e9, "tflops_fp32_equiv": a.chains * fl / (ms * 1e-3) / 1e12,
                      "workspace_gb": prep.workspace(a.chains).numel() / 2**30, "logp0": float(logp[0])}))


if __name__ == "__main__":
    main()
