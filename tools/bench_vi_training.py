"""Time the Bayes-by-Backprop trainer at the reference's configuration (Neural_network/VI/config.py: 1-10-10-1 tanh, 10 000 epochs,
num_ens = 10, Adam 1e-2) on the bundled data: CUDA engine (one CUDA-graph replay per epoch) next to the torch-CPU restatement
of the reference loop (bounded sample of epochs, one thread -- the loop is dispatch-bound)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]

import torch  # noqa: E402

from vihmc import synth, vi  # noqa: E402
from vihmc.spec import LogProbSpec, MLPArch  # noqa: E402


def main():
    epochs, num_ens = 10_000, 10
    x, y, xv, yv = synth.bnn_data()
    arch = MLPArch(in_dim=1, widths=(10, 10), out_dim=1, act="tanh", last_bias=True)
    mk = lambda a, b: LogProbSpec(arch=arch, x=a, y=b, loss="NLL", tau_out=0.05 ** 2, prior_sigma_scalar=1.0)
    vi.train_bbb(mk(x, y), mk(xv, yv), epochs=50, num_ens=num_ens, seed=1)       # warm-up: library load, kernel attributes
    for graph in (True, False):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = vi.train_bbb(mk(x, y), mk(xv, yv), epochs=epochs, num_ens=num_ens, lr_start=1e-2, lr_patience=5000, seed=1, use_graph=graph)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        h = res.history.numpy()
        print(json.dumps({"workload": "BNN Bayes-by-Backprop, 10000 epochs x num_ens 10, host tensors in / artefacts out",
                          "cuda_graph": graph, "seconds": dt, "us_per_epoch": 1e6 * dt / epochs,
                          "train_loss_first_last": [float(h[0, 0]), float(h[-1, 0])], "valid_loss_best": float(h[:, 1].min()),
                          "final_lr": float(h[-1, 2])}))
    # CPU restatement of the reference loop, bounded sample
    from oracle import vi_bbb as ovi
    torch.set_num_threads(1)
    n = 200
    mu0, rho0 = vi.init_posterior(141, vi.DEFAULT_PRIORS, 1)
    eps = torch.randn(n, num_ens, 141)
    t0 = time.perf_counter()
    ovi.train(x, y, xv, yv, (10, 10), "tanh", mu0, rho0, eps, 0.05 ** 2, 0.0, 1.0, 1e-2, 5000)
    dt = time.perf_counter() - t0
    print(json.dumps({"workload": "torch-CPU restatement of the reference loop (kind: port), 1 thread", "epochs_timed": n,
                      "us_per_epoch": 1e6 * dt / n, "extrapolated_seconds_for_10000_epochs": dt / n * epochs}))


if __name__ == "__main__":
    main()
