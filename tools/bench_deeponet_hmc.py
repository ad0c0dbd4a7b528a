#!/usr/bin/env python
"""End-to-end DeepONet HMC at the BASELINE.json sizes (configs[2] full HMC incl. the M=2 split integrator,
configs[3] VI-HMC with d = 10 % of D) through the public API: samplers.sample(spec, q0_host, ...).

  python tools/bench_deeponet_hmc.py --chains 64 --samples 3 [--mode full|split|vi]

N=1000 branch functions, P=10201 trunk points, D=172401, L=7, eps=1e-4 (Operator_network/*/config*.py).
Prints one JSON line: chain-grad-evals/s end to end (host tensors in, host samples out)."""
import argparse, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import numpy as np
import torch
from vihmc import samplers, synth
from vihmc.spec import DeepONetArch, LogProbSpec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--samples", type=int, default=3)
    ap.add_argument("--mode", default="full", choices=["full", "split", "vi"])
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--out-scale", type=float, default=1.0, help="synth.burgers_like(out_scale=...): 0.15 = targets in the +-0.2 range")
    ap.add_argument("--q0-noise", type=float, default=0.001)
    ap.add_argument("--eps", type=float, default=3e-5,
                    help="step size.  The reference config's 1e-4 is tuned for its trained net; on the synthetic teacher problem the "
                         "leapfrog is unstable above ~3e-5 (non-finite H1 => every proposal rejected, as hamiltorch would)")
    a = ap.parse_args()
    arch = DeepONetArch()
    rs = np.random.RandomState(0)
    P = 101 * 101
    # teacher-generated Burgers-shaped data: chains start next to theta*; the default step size 3e-5 is stable there (1e-4 is not, see --eps)
    x1, x2, y, theta = synth.burgers_like(arch, n_train=a.n, n_t=101, n_x=101, seed=0, out_scale=a.out_scale)
    kw = dict(arch=arch, x2=x2, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1)
    L, eps = 7, a.eps
    if a.mode == "vi":
        mu, sigma, ind = synth.deeponet_vi_artifacts(theta, 0.10, seed=1)
        spec = LogProbSpec(x=x1, y=y, frozen=mu, sens_ind=ind, **kw)
        q0 = mu[ind][None].repeat(a.chains, 1)
        integ, evals = samplers.Integrator.IMPLICIT, a.samples * (L + 1)
    elif a.mode == "split":
        h = a.n // 2
        spec = [LogProbSpec(x=x1[i * h:(i + 1) * h], y=y[i * h:(i + 1) * h], prior_scale=2.0, **kw) for i in range(2)]
        q0 = theta[None].repeat(a.chains, 1)
        integ, evals = samplers.Integrator.SPLITTING, a.samples * L * 4 / 2   # 4 half-data gradient evals per step
    else:
        spec = LogProbSpec(x=x1, y=y, **kw)
        q0 = theta[None].repeat(a.chains, 1)
        integ, evals = samplers.Integrator.IMPLICIT, a.samples * (L + 1)
    q0 = q0 + a.q0_noise * torch.from_numpy(rs.randn(*q0.shape).astype(np.float32))
    samplers.sample(spec, q0, num_samples=1, num_steps_per_sample=1, step_size=eps, integrator=integ)   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = samplers.sample(spec, q0, num_samples=a.samples, num_steps_per_sample=L, step_size=eps, integrator=integ,
                          return_result=True, seed=1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({"workload": f"deeponet {a.mode} HMC N={a.n} P={P} D={arch.num_params} d={q0.shape[1]} chains={a.chains} "
                                  f"samples={a.samples} L={L} eps={eps}",
                      "seconds_e2e": dt, "chain_grad_evals_per_s_e2e": a.chains * evals / dt,
                      "acceptance_rate": res.acceptance_rate, "samples_shape": list(res.samples.shape),
                      "H0_mean": float(res.hamiltonians[:, :, 0].mean()), "dH_abs_mean": float((res.hamiltonians[:, :, 0] - res.hamiltonians[:, :, 1]).abs().mean())}))


if __name__ == "__main__":
    main()
