#!/usr/bin/env python
"""BASELINE.json configs[3]: DeepONet VI-HMC (d = 10 % of D = 172 401), chains sharded over the GPUs of one box, NCCL used
only AFTER sampling: gather of (thinned) samples to rank 0 and an all-gather of per-half-chain moments for the global split-R-hat.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/bench_deeponet_sharded.py \
      --chains 4096 --samples 5            (G = 1: python tools/bench_deeponet_sharded.py --chains 512)

Timing: barrier + synchronize on both sides, CUDA events on the sampling stream, MAX over ranks.  Prints one JSON line
on rank 0: chain-grad-evals/s of the whole job, the gather time, and the R-hat summary (with so few draws it only
demonstrates the collective path -- the chains are still in burn-in)."""
import argparse, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "vi-hmc_b200")]
import numpy as np
import torch
import torch.distributed as dist
from vihmc import dist as vd, engine, synth
from vihmc.spec import DeepONetArch, LogProbSpec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=4096, help="total over all ranks")
    ap.add_argument("--samples", type=int, default=5)
    ap.add_argument("--eps", type=float, default=3e-5)
    ap.add_argument("--gather-every", type=int, default=1, help="thinning of the gathered samples")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    arch = DeepONetArch()
    x1, x2, y, theta = synth.burgers_like(arch, n_train=1000, n_t=101, n_x=101, seed=0)
    mu, sigma, ind = synth.deeponet_vi_artifacts(theta, 0.10, seed=1)
    spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1, frozen=mu, sens_ind=ind)
    L = 7
    chain0, n_local = vd.shard_chains(a.chains)
    # q0 = mu[ind] + sigma[ind] * z, z keyed by the GLOBAL chain id (only this rank's rows are materialised)
    q0 = torch.stack([mu[ind] + sigma[ind] * torch.from_numpy(np.random.RandomState(1000 + chain0 + c).randn(len(ind)).astype(np.float32))
                      for c in range(n_local)])
    prep = engine.prepare(spec, dev)
    kw = dict(num_samples=a.samples, num_steps=L, step_size=a.eps, burn=0, seed=1, chain_offset=chain0, to_host=False)
    engine.run_sampler([prep], q0[:min(n_local, 8)], num_samples=1, num_steps=1, step_size=a.eps, to_host=False)   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = engine.run_sampler([prep], q0, **kw)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # collectives after sampling
    t0 = time.perf_counter()
    rhat = vd.global_split_rhat(res.samples) if res.samples.shape[0] >= 4 else None
    gathered = vd.gather_chains(res.samples[::a.gather_every].contiguous(), a.chains)
    acc = vd.gather_chains(res.accepted, a.chains)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_coll = time.perf_counter() - t0
    if rank == 0:
        evals = a.chains * a.samples * (L + 1)
        print(json.dumps({"workload": f"deeponet VI-HMC cfg4: N=1000 P=10201 D={arch.num_params} d={len(ind)} chains={a.chains} "
                                      f"samples={a.samples} L={L} eps={a.eps}",
                          "n_gpus": world, "chains_per_gpu": n_local, "sampling_ms_max_over_ranks": ms,
                          "chain_grad_evals_per_s": evals / (ms * 1e-3), "gather_and_rhat_s": t_coll,
                          "gathered_samples_shape": list(gathered.shape), "gathered_GB": gathered.numel() * 4 / 1e9,
                          "acceptance_rate": float(acc.float().mean()),
                          "rhat_max": None if rhat is None else float(rhat.max()), "rhat_median": None if rhat is None else float(rhat.median())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
