"""Diagnostic (not collected by pytest): forward output of the engine at BASELINE size against the fp64 oracle forward --
coherent (mean signed) and rms error of S = model(q), for the path selected by the environment."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "vi-hmc_b200"), HERE]
from oracle import closures as oc  # noqa: E402
from vihmc import engine, synth  # noqa: E402
from vihmc.spec import DeepONetArch, LogProbSpec  # noqa: E402


def main():
    g = np.load(os.path.join(HERE, "golden", "deeponet_fullsize_logp_grad.npz"))
    arch = DeepONetArch()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    x1, x2, y, theta = synth.burgers_like(arch, n_train=1000, n_t=101, n_x=101, seed=0)
    x1, y = x1[:n], y[:n]
    spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1)
    q = torch.from_numpy(g["full/q"][:1])
    pred = engine.predict(spec, q)[0].double().cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    cl64 = oc.DeepONetLogProb(x1=x1.unsqueeze(1), x2=x2.unsqueeze(0), y=y, dtype=torch.float64)
    ref = cl64.forward(q[0].double()).detach()
    cl32 = oc.DeepONetLogProb(x1=x1.unsqueeze(1), x2=x2.unsqueeze(0), y=y)
    r32 = cl32.forward(q[0]).detach().double()
    print("env", {k: v for k, v in os.environ.items() if k.startswith("VIHMC")})
    for name, out in (("engine", pred), ("torch fp32 (the reference's arithmetic)", r32)):
        e = out - ref
        print(f"{name}: mean S {float(ref.mean()):+.4f}; coherent abs error {float(e.mean()):+.3e} (relative to |mean S| {float(e.mean() / ref.mean().abs()):+.2e}); "
              f"rms {float(e.pow(2).mean().sqrt()):.3e}; max {float(e.abs().max()):.3e}; regression slope of err on S {float((e * ref).sum() / (ref * ref).sum()):+.3e}")


if __name__ == "__main__":
    main()
