#!/usr/bin/env python
"""BASELINE.json configs[4]: wide BNN (1-512-512-512-512-1 tanh, D = 789 505) full-parameter HMC on a 100k-point
synthetic set, training rows sharded over the ranks, gradient all-reduce once per evaluation.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_wide_bnn.py --chains 8
  python tools/bench_wide_bnn.py --chains 8           (G = 1)

Prints one JSON line on rank 0: chain-grad-evals/s, FP32-equivalent TFLOP/s (472.4 GFLOP per unit at N = 100k,
SURVEY.md 8(d)), and a checksum of the final samples (identical for any G up to fp32 summation order)."""
import argparse, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import numpy as np
import torch
import torch.distributed as dist
from vihmc import dist as vd, synth
from vihmc.spec import LogProbSpec, MLPArch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=8)
    ap.add_argument("--samples", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--eps", type=float, default=2e-6)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    arch = MLPArch(in_dim=1, widths=(512, 512, 512, 512), out_dim=1, act="tanh", last_bias=True)
    x, y = synth.wide_bnn_data(a.n, seed=0)
    spec = LogProbSpec(arch=arch, x=x, y=y, loss="NLL", tau_out=0.0025, prior_sigma_scalar=1.0)
    q0 = synth.default_linear_init(arch, seed=0)[None].repeat(a.chains, 1)
    q0 = q0 + 1e-3 * torch.from_numpy(np.random.RandomState(1).randn(*q0.shape).astype(np.float32))
    local_spec = vd.shard_spec_rows(spec)
    from vihmc import engine
    local_prep = engine.prepare(local_spec)      # prepared once: the warm-up call captures the step graphs, the timed call replays them
    vd.sample_data_sharded(local_prep, q0, 1, 2, a.eps, seed=0)   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = vd.sample_data_sharded(local_prep, q0, a.samples, a.steps, a.eps, seed=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    evals = a.chains * a.samples * (a.steps + 1)
    dims = [1, 512, 512, 512, 512, 1]
    macs = sum(dims[i] * dims[i + 1] for i in range(5))
    flop = 2 * a.n * (3 * macs - dims[0] * dims[1])
    if rank == 0:
        print(json.dumps({"workload": f"wide BNN 4x512 D={arch.num_params} N={a.n} chains={a.chains} samples={a.samples} L={a.steps}",
                          "n_gpus": world, "rows_per_gpu": local_spec.N, "seconds": dt, "chain_grad_evals_per_s": evals / dt,
                          "gflop_per_unit": flop / 1e9, "tflops_fp32_equiv": evals * flop / dt / 1e12,
                          "allreduce_bytes_per_eval": a.chains * arch.num_params * 4,
                          "acceptance": float(out["accepted"].float().mean()),
                          "dH_abs_mean": float((out["hamiltonians"][..., 0] - out["hamiltonians"][..., 1]).abs().mean()),
                          "checksum": float(out["samples"][-1].double().sum())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
