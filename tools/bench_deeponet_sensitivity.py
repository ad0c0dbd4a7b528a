"""Time vihmc_deeponet_sensitivity at the shipped size (Operator_network/VI/config_sens.py: 100-wide 9/9 DeepONet, N_valid = 1000
functions) on the full 101 x 101 trunk grid and on a p = 100 subset (the reference's cfg.p).  CUDA events, inputs resident."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]

import torch  # noqa: E402

from vihmc import sensitivity as vs, synth  # noqa: E402
from vihmc.spec import DeepONetArch  # noqa: E402


def main():
    arch = DeepONetArch()
    x1, x2, y, theta = synth.burgers_like(arch, n_train=1000, n_t=101, n_x=101, seed=0)
    sigma = 0.001 + 0.01 * torch.rand(arch.num_params, generator=torch.Generator().manual_seed(5))
    dev = torch.device("cuda", 0)
    w, sg = theta.to(dev), sigma.to(dev)
    for label, trunk in (("full grid P=10201", x2), ("subset P=100", x2[torch.randperm(x2.shape[0], generator=torch.Generator().manual_seed(1))[:100]])):
        vs._deeponet_batch_scores(arch, x1, trunk, w, sg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            s = vs._deeponet_batch_scores(arch, x1, trunk, w, sg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        P = trunk.shape[0]
        print(json.dumps({"workload": f"deeponet sensitivity, N=1000, {label}, D=172401", "ms_per_call_incl_h2d": ms,
                          "jacobian_entries_never_formed": 1000 * P * arch.num_params,
                          "scores_min_max": [float(s.min()), float(s.max())]}))


if __name__ == "__main__":
    main()
