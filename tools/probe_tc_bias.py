"""Probe: signed relative error of the DeepONet forward output (one K=100 GEMM chain per layer + head)
against the fp64 oracle, for the tensor-core and the FP32-SIMT GEMM paths (VIHMC_DENSE_SIMT=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from oracle import closures as oc
from vihmc import engine, synth
from vihmc.spec import DeepONetArch, LogProbSpec
arch = DeepONetArch()
x1, x2, y, theta = synth.burgers_like(arch, n_train=200, n_t=30, n_x=50, seed=5)
spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1)
closure = oc.DeepONetLogProb(x1=x1.unsqueeze(1), x2=x2.unsqueeze(0), y=y, dtype=torch.float64)
q = theta[None]
O = engine.predict(spec, q).cpu().double()[0]
ref = closure.forward(theta.double()).detach()
err = (O - ref)
big = ref.abs() > 0.1 * ref.abs().mean()
print("mode", "SIMT" if os.environ.get("VIHMC_DENSE_SIMT") == "1" else "TC",
      "| mean|ref| %.3e" % ref.abs().mean().item(),
      "| mean signed rel err (err*sign(ref)/|ref|) %.3e" % ((err * ref.sign() / ref.abs())[big]).mean().item(),
      "| rms rel err %.3e" % ((err / ref.abs())[big] ** 2).mean().sqrt().item(),
      "| sum err %.4f" % err.sum().item())
