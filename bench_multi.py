"""Multi-GPU legs of bench.py (imported by it under torchrun, N > 1): the two BASELINE configs whose data path communicates.

cfg4  BASELINE configs[3]: DeepONet VI-HMC (d = 10 % of D = 172 401), chains sharded over the GPUs (weak: CFG4_CHAINS_PER_GPU
      each); the e2e clock covers sampling, the NCCL gather of every stored draw to rank 0 and the all-gather of the
      per-half-chain moments for the global split-R-hat -- the collectives Operator_network/VI_HMC would need to pool its chains.
cfg5  BASELINE configs[4]: wide BNN 4x512, 100 000 rows sharded over the GPUs, every rank holds all 8 chains, ONE all-reduce of the
      [C, d + 1] buffer (gradient with the log-posterior packed behind it) per gradient evaluation (strong scaling: the total
      work is fixed).  all_reduce_ms is the CUDA-event time of the collectives alone, measured in a second pass.
Timing: barrier + synchronize on both sides, CUDA events / wall clock, MAX over ranks."""
from __future__ import annotations

import time

CFG4_CHAINS_PER_GPU, CFG4_SAMPLES, CFG4_L, CFG4_EPS = 512, 5, 7, 1e-4   # BASELINE configs[3]: 4096 chains over 8 GPUs
CFG5_CHAINS, CFG5_ROWS, CFG5_SAMPLES, CFG5_L, CFG5_EPS = 8, 100_000, 4, 4, 2e-6
CFG5_GFLOP = 472.4          # per chain-grad-eval over all 100k rows


def _max_over_ranks(dist, dev, value):
    import torch

    if dist is None or not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(dist):
    if dist is not None and dist.is_initialized():
        dist.barrier()


def cfg4_leg(dev, rank, world, dist):
    import numpy as np
    import torch
    from vihmc import dist as vd, engine, synth
    from vihmc.spec import DeepONetArch, LogProbSpec

    arch = DeepONetArch()
    x1, x2, y, theta = synth.burgers_like(arch, n_train=1000, n_t=101, n_x=101, seed=0, out_scale=0.39, trunk_scale=0.39)
    mu, sigma, ind = synth.deeponet_vi_artifacts(theta, 0.10, seed=1)
    mu = theta + 0.001 * (mu - theta)                 # fitted means: next to the teacher (the synthetic artefacts sit 0.01 away)
    spec = LogProbSpec(arch=arch, x=x1, x2=x2, y=y, loss="NLL", tau_out=1.0, prior_sigma_scalar=0.1, frozen=mu, sens_ind=ind)
    total = CFG4_CHAINS_PER_GPU * world
    chain0, n_local = vd.shard_chains(total)
    q0 = torch.stack([mu[ind] + 0.1 * sigma[ind] * torch.from_numpy(np.random.RandomState(7000 + chain0 + c).randn(len(ind)).astype(np.float32))
                      for c in range(n_local)])
    prep = engine.prepare(spec, dev)
    warm = engine.run_sampler([prep], q0[:8], num_samples=4, num_steps=1, step_size=CFG4_EPS, to_host=False)   # warm-up: kernels ...
    vd.global_split_rhat(warm.samples)                                    # ... and the collectives (NCCL sets up its channels on first use)
    vd.gather_chains(warm.samples, 8 * world)
    vd.gather_chains(warm.accepted, 8 * world)
    del warm
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    res = engine.run_sampler([prep], q0, num_samples=CFG4_SAMPLES, num_steps=CFG4_L, step_size=CFG4_EPS, burn=0, seed=1,
                             chain_offset=chain0, to_host=False)
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(dist, dev, e0.elapsed_time(e1))
    tc0 = time.perf_counter()
    rhat = vd.global_split_rhat(res.samples)
    gathered = vd.gather_chains(res.samples, total)
    acc = vd.gather_chains(res.accepted, total)
    torch.cuda.synchronize()
    dist.barrier()
    t_coll = _max_over_ranks(dist, dev, time.perf_counter() - tc0)
    e2e_s = _max_over_ranks(dist, dev, time.perf_counter() - t0)
    if rank != 0:
        return None
    evals = total * CFG4_SAMPLES * (CFG4_L + 1)
    return {"workload": f"deeponet_vi_hmc cfg4: N=1000 x P=10201, D=172401, d={len(ind)}, L={CFG4_L}, eps={CFG4_EPS}, "
                        f"{CFG4_CHAINS_PER_GPU} chains/GPU x {world} GPUs, {CFG4_SAMPLES} iterations",
            "scaling": "weak", "value": evals / (ms * 1e-3), "unit": "chain-grad-evals/s", "sampling_ms_max_over_ranks": ms,
            "e2e": {"value": evals / e2e_s, "unit": "chain-grad-evals/s", "seconds": e2e_s,
                    "includes": "sampling + NCCL gather of all stored draws to rank 0 + all-gather of half-chain moments + global split-R-hat"},
            "collectives_seconds": t_coll, "gathered_bytes": int(gathered.numel() * 4),
            "acceptance_rate": float(acc.float().mean()), "rhat_max_over_5_draws": float(rhat.max())}


def cfg5_leg(dev, rank, world, dist):
    import torch
    from vihmc import dist as vd, engine, synth
    from vihmc.spec import LogProbSpec, MLPArch

    arch = MLPArch(in_dim=1, widths=(512, 512, 512, 512), out_dim=1, act="tanh", last_bias=True)
    x, y = synth.wide_bnn_data(n=CFG5_ROWS, seed=0)
    spec = LogProbSpec(arch=arch, x=x, y=y, loss="NLL", tau_out=0.0025, prior_sigma_scalar=1.0)
    local = engine.prepare(vd.shard_spec_rows(spec), dev)   # prepared once: the warm-up call captures the step graphs, the timed call replays them
    q0 = synth.default_linear_init(arch, seed=0).unsqueeze(0).repeat(CFG5_CHAINS, 1)
    q0 = q0 + 1e-3 * torch.randn(q0.shape, generator=torch.Generator().manual_seed(3))
    kw = dict(num_samples=CFG5_SAMPLES, num_steps=CFG5_L, step_size=CFG5_EPS, burn=0, seed=5)
    vd.sample_data_sharded(local, q0, num_samples=1, num_steps=2, step_size=CFG5_EPS)      # warm-up (captures the three step graphs)
    torch.cuda.synchronize()
    _barrier(dist)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = vd.sample_data_sharded(local, q0, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(dist, dev, e0.elapsed_time(e1))
    n_evals = CFG5_SAMPLES * (CFG5_L + 1)
    # the collectives alone: the same number of all-reduces of the same buffer, back to back
    buf = torch.zeros((CFG5_CHAINS, arch.num_params + 1), device=dev)
    ar_ms = 0.0
    if world > 1:
        dist.all_reduce(buf)
        torch.cuda.synchronize()
        _barrier(dist)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(n_evals):
            dist.all_reduce(buf)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = _max_over_ranks(dist, dev, a0.elapsed_time(a1))
    if rank != 0:
        return None
    evals = CFG5_CHAINS * n_evals
    value = evals / (ms * 1e-3)
    return {"workload": f"wide_bnn cfg5: 1-512-512-512-512-1 tanh, D=d={arch.num_params}, {CFG5_ROWS} rows sharded over {world} GPUs, "
                        f"{CFG5_CHAINS} chains on every GPU, L={CFG5_L}, {CFG5_SAMPLES} iterations, one all-reduce of [C, d+1] per evaluation",
            "scaling": "strong", "value": value, "unit": "chain-grad-evals/s", "ms_max_over_ranks": ms,
            "tflops_fp32_equivalent": value * CFG5_GFLOP / 1e3, "all_reduce_bytes_per_eval": int(buf.numel() * 4),
            "all_reduce_ms_total": ar_ms, "all_reduce_share_of_time": ar_ms / ms,
            "cuda_graphs": bool(out.get("graphs", False)),
            "acceptance_rate": float(out["accepted"].float().mean())}
