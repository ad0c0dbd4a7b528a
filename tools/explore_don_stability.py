"""Which synthetic Burgers-shaped teachers make the reference's DeepONet step size (1e-4, L = 7) a stable leapfrog?  Acceptance rate and
energy error of a few HMC iterations at full size for a grid of feature scales (vihmc.synth.burgers_like(out_scale, trunk_scale))."""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import numpy as np
import torch
from vihmc import engine, synth
from vihmc.spec import DeepONetArch, LogProbSpec

arch = DeepONetArch()
for bs, ts in [(1.0, 1.0), (0.15, 1.0), (0.39, 0.39), (0.15, 0.15), (0.15, 0.05), (0.05, 0.05), (0.5, 0.03)]:
    x1, x2, y, theta = synth.burgers_like(arch, n_train=1000, n_t=101, n_x=101, seed=0, out_scale=bs, trunk_scale=ts)
    spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1)
    prep = engine.prepare(spec)
    q0 = theta[None].repeat(32, 1) + 0.001 * torch.from_numpy(np.random.RandomState(0).randn(32, arch.num_params).astype(np.float32))
    for eps in (1e-4, 3e-5):
        res = engine.run_sampler([prep], q0, 4, 7, eps, seed=1, to_host=False)
        dH = (res.hamiltonians[..., 1] - res.hamiltonians[..., 0])
        print(json.dumps({"branch_scale": bs, "trunk_scale": ts, "eps": eps, "y_rms": float(y.pow(2).mean().sqrt()), "accept": float(res.accepted.float().mean()),
                          "dH_median": float(dH.median()), "dH_max": float(dH.max())}), flush=True)
    del prep
    torch.cuda.empty_cache()
