"""Mean signed error of the dense path's tanh implementations on the GPU (vihmc_debug_tanh) against fp64, for Gaussian
pre-activations of several widths.  'toward |.|' = mean of err * sign(x) / |tanh x|: negative = shrinks the activations."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import torch
from vihmc import _lib

lib = _lib.load()
names = {0: "ex2.approx/rcp.approx form (tanh_sel)", 1: "CUDA tanhf", 2: "tanh_acc2 (exact forward pass)"}
for sig in (0.1, 0.3, 1.0, 2.0):
    g = torch.Generator(device="cuda").manual_seed(int(sig * 10))
    x = (sig * torch.randn(1 << 24, generator=g, device="cuda")).float()
    ref = torch.tanh(x.double())
    for kind in (0, 1, 2):
        y = torch.empty_like(x)
        _lib.check(lib.vihmc_debug_tanh(kind, x.data_ptr(), y.data_ptr(), x.numel(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        err = y.double() - ref
        rel = err * torch.sign(ref) / ref.abs().clamp_min(1e-9)
        ulp = (torch.nextafter(ref.float().abs(), torch.full_like(x, 2.0)) - ref.float().abs()).double()
        print(json.dumps({"sigma": sig, "impl": names[kind], "mean_rel_err_toward_abs": float(rel.mean()), "mean_abs_err_signed": float((err * torch.sign(ref)).mean()),
                          "rms_rel": float(rel.pow(2).mean().sqrt()), "max_ulp": float((err.abs() / ulp).max())}))
