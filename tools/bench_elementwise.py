#!/usr/bin/env python
"""HBM roofline of the leapfrog building blocks at the DeepONet shape (BASELINE.json configs[2]: 256 chains x d = 172 401).

  python tools/bench_elementwise.py [--chains 256] [--d 172401] [--reps 20]

Every array is chains * d * 4 B = 176 MB > the 126 MB L2, so each launch streams from HBM.  Algorithmic bytes per
chain and coordinate: leapfrog update 20 B (read q, p, g; write q, p), kick only 12 B, momentum draw
4 B, Metropolis select 12 B (accepted, stored: read proposal, write current + fallback + stored = 16 B; counted as
written), VI scatter 8 B, gather 8 B.  Prints one JSON line per kernel: GB/s and fraction of the measured copy peak
(MEASURED_PEAKS.json hbm_gbs, fallback 6547.8)."""
import argparse, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import numpy as np
import torch
from vihmc import engine


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=256)
    ap.add_argument("--d", type=int, default=172401)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    C, d = a.chains, a.d
    dev = torch.device("cuda:0")
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6547.8
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(C, d, device=dev, generator=g)
    p = torch.randn(C, d, device=dev, generator=g)
    gr = torch.randn(C, d, device=dev, generator=g)
    qf = q.clone()
    stored = torch.empty_like(q)
    H0 = torch.zeros(C, device=dev)
    H1 = torch.zeros(C, device=dev) - 1.0          # every chain accepts
    u = torch.full((C,), 0.5, device=dev)
    acc = torch.empty(C, dtype=torch.uint8, device=dev)
    D = d
    ind = np.sort(np.random.RandomState(0).choice(D, D // 10, replace=False)).astype(np.int64)
    frozen = torch.randn(D, device=dev, generator=g)
    qs = torch.randn(C, len(ind), device=dev, generator=g)
    lib_cases = [
        ("leapfrog_update_kernel (kick + drift + kinetic energy)", 20, lambda: engine.leapfrog_update(q, p, gr, 1e-6, 1.0, 1.0, want_ke=True)),
        ("leapfrog_update_kernel (kick only)", 12, lambda: engine.leapfrog_update(q, p, gr, 1e-6, 0.5, 0.0)),
        ("momentum_philox_kernel", 4, lambda: engine.momentum_philox(1, 0, 0, C, d, dev)),
        ("mh_accept_kernel (accept + store)", 16, lambda: engine.mh_accept(H0, H1, u, q, p, qf, stored, acc)),
        ("copy (torch, the peak's own definition)", 8, lambda: stored.copy_(q)),
    ]
    for name, bytes_per, fn in lib_cases:
        ms = timed(fn, a.reps)
        gbs = C * d * bytes_per / (ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "chains": C, "d": d, "ms": ms, "algorithmic_bytes_per_coord": bytes_per, "GB/s": gbs,
                          "peak_GB/s": peak, "frac": gbs / peak}))
    ms = timed(lambda: engine.scatter_vi(frozen, ind, qs), a.reps)
    gbs = C * D * 4 / (ms * 1e-3) / 1e9      # written bytes: the full [C, D] weight matrix
    print(json.dumps({"kernel": "scatter_vi (fill + put, d = D/10)", "chains": C, "D": D, "ms": ms, "GB/s_written": gbs, "peak_GB/s": peak,
                      "frac": gbs / peak, "note": "includes the torch.empty of W and the index upload of the python wrapper"}))


if __name__ == "__main__":
    main()
