#!/usr/bin/env python
"""bench.py -- chain-grad-evals/sec of the VI-HMC hot path on B200 (BASELINE.json metric).

Workload (N=1): BASELINE.json configs[1] -- Neural_network/VI_HMC: BNN 1-10-10-1 tanh on the bundled
20-point set, HMC over the VI-selected subset (d=40 of D=141, synthetic artefacts of SURVEY.md 8(d)),
step_size 5e-4, L=196, NLL tau_out=0.0025, 1024 chains per GPU (weak scaling: chains sharded, no
data-path collective).  One STEP = one HMC iteration (momentum draw, H0, L leapfrog steps, H1,
Metropolis test, store) of all chains = L+1 = 197 log-posterior gradient evaluations per chain.

  python bench.py [--gpus N] [--steps K] [--warmup W]         our CUDA engine
  python bench.py --impl reference [--steps K] [--warmup W]   the reference's CPU path (oracle port)

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "vi-hmc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC, UNIT = "chain-grad-evals/sec", "chain-grad-evals/s"
_REAL_STDOUT_FD = None               # set when fd 1 is redirected (multi-GPU runs, see run_ours)
CHAINS_PER_GPU = 1024
D_SAMPLED, STEP_SIZE, L_STEPS, TAU_OUT = 40, 5e-4, 196, 0.0025
FLOP_PER_GRAD_EVAL = 14_000          # 2*[3*N*sum(in*out) - N*in_1*out_1], N=20, 1-10-10-1
REF_EVALS_PER_STEP = 16              # reference arm: bounded sample = 16 grad-evals per chain per step
WORKLOAD = "bnn_vi_hmc cfg2: 1-10-10-1 tanh, N=20, d=40 of D=141, L=196, eps=5e-4, NLL v=0.0025, 1024 chains/GPU"


def build_spec():
    import numpy as np
    import torch
    from vihmc import synth
    from vihmc.spec import LogProbSpec, sliced_prior_sigma

    x, y, _, _ = synth.bnn_data()
    arch = synth.bnn_arch()
    mu, sigma, ind = synth.bnn_vi_artifacts(arch.num_params, D_SAMPLED, seed=1)
    sig = sliced_prior_sigma(D_SAMPLED, arch.tensor_numels(), [1.0] * 6)
    spec = LogProbSpec(arch=arch, x=x, y=y, loss="NLL", tau_out=TAU_OUT, prior_sigma=torch.from_numpy(sig.astype(np.float32)),
                       frozen=mu, sens_ind=ind, vi_sigma=sigma)
    return spec, mu, sigma, ind


def initial_states(mu, sigma, ind, chains, chain0):
    """q0 = mu[ind] + sigma[ind] * eps_c, eps_c from numpy keyed by the global chain id."""
    import numpy as np
    import torch

    out = np.empty((chains, len(ind)), dtype=np.float32)
    m, s = mu.numpy()[ind], sigma.numpy()[ind]
    for c in range(chains):
        out[c] = m + s * np.random.RandomState(1000 + chain0 + c).randn(len(ind))
    return torch.from_numpy(out)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (NVML; falls back to nvidia-smi)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """NVML is initialised and the first sample taken synchronously in __enter__ (nvmlInit can take longer than the whole
    timed region on a fresh box); the thread then samples every 2 ms until __exit__, which takes a last sample itself."""

    _NAMES = None

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.stop = index, [], set(), threading.Event()
        self.max_mhz = None
        self.nv = self.h = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, nm in self._names.items():
            if mask & bit:
                self.reasons.add(nm)

    def _run(self):
        try:
            while not self.stop.is_set():
                self._sample()
                time.sleep(0.002)
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"clock_sampling_failed:{type(e).__name__}")

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self._names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                           nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                           nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                           nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            self._sample()
            self.thread.start()
        except Exception as e:  # pragma: no cover
            self.reasons.add(f"clock_sampling_failed:{type(e).__name__}")
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.thread.is_alive():
            self.thread.join(timeout=2)
        try:
            if self.nv is not None:
                self._sample()
        except Exception:  # pragma: no cover
            pass

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (torch-eager closure + restated hamiltorch leapfrog)
# ------------------------------------------------------------------------------------------------
def _ref_worker(idx, steps, warmup, barrier, out):
    import torch
    torch.set_num_threads(1)
    from oracle import closures as oc
    from oracle import hamiltorch_restated as hr

    spec, mu, sigma, ind = build_spec()
    closure = oc.BnnLogProb(x=spec.x, y=spec.y, widths=(10, 10), act="tanh", loss="NLL", tau_out=TAU_OUT,
                            prior=("sliced", [1.0] * 6), frozen=mu, sens_ind=ind)
    q = initial_states(mu, sigma, ind, 1, idx)[0]
    g = torch.Generator().manual_seed(idx)
    for _ in range(warmup):
        q, _ = hr.leapfrog(q, torch.randn(q.shape, generator=g), closure, REF_EVALS_PER_STEP - 1, STEP_SIZE)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        q, _ = hr.leapfrog(q, torch.randn(q.shape, generator=g), closure, REF_EVALS_PER_STEP - 1, STEP_SIZE)
    out[idx] = time.perf_counter() - t0
    barrier.wait()


def run_reference(steps, warmup, n_gpus):
    """One chain per host core (torch threads = 1), all cores in parallel -- the better CPU mode for this
    2-3 ms, dispatch-bound closure."""
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    import multiprocessing as mp

    cores = len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(cores + 1)
    out = ctx.Array("d", cores)
    procs = [ctx.Process(target=_ref_worker, args=(i, steps, warmup, barrier, out)) for i in range(cores)]
    for p in procs:
        p.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    wall = time.perf_counter() - t0
    for p in procs:
        p.join()
    evals = cores * steps * REF_EVALS_PER_STEP
    value = evals / wall
    sample = (f"{cores} processes x 1 chain x {steps} steps x {REF_EVALS_PER_STEP} grad-evals "
              f"(leapfrog segments of the cfg2 trajectory), torch-eager closure + autograd.grad")
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 1e3 * wall / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "step": f"{REF_EVALS_PER_STEP} grad-evals per chain, one chain per core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def cpu_baseline_leg():
    """Bounded CPU sample for our arm's JSON line: run the reference arm in a fresh process (no CUDA context)."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "60", "--warmup", "3"],
                           capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
        line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e}"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(steps, warmup, n_gpus, skip_cpu=False):
    import torch
    import torch.distributed as dist
    from vihmc import engine, samplers

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # the contract is ONE line on stdout, but NCCL writes its "NCCL version ..." banner to file descriptor 1 when the
        # communicator is created: point fd 1 at stderr for the run and keep the real stdout for the JSON line
        global _REAL_STDOUT_FD
        sys.stdout.flush()
        _REAL_STDOUT_FD = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    warmup = max(warmup, 3)
    spec, mu, sigma, ind = build_spec()
    chain0 = rank * CHAINS_PER_GPU
    q0_host = initial_states(mu, sigma, ind, CHAINS_PER_GPU, chain0)
    prep = engine.prepare(spec, dev)
    q0 = q0_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def launch(n_samples, seed):
        return engine.run_sampler([prep], q0, n_samples, L_STEPS, STEP_SIZE, burn=0, seed=seed, chain_offset=chain0,
                                  diagnostics=True, to_host=False)

    # ---- device-timed: inputs resident in HBM, one persistent launch of `steps` iterations ----
    launch(warmup, seed=1)
    torch.cuda.synchronize()
    # the sampler allocates its result tensors with torch.empty: put blocks of those sizes into torch's caching allocator
    # now, so that no cudaMalloc (a host-side, millisecond-scale call) lands between the two timing events
    warm = [torch.empty((steps, CHAINS_PER_GPU, D_SAMPLED), device=dev), torch.empty((steps, CHAINS_PER_GPU, 2), device=dev),
            torch.empty((steps, CHAINS_PER_GPU), device=dev), torch.empty((steps, CHAINS_PER_GPU), dtype=torch.uint8, device=dev),
            torch.empty(CHAINS_PER_GPU, device=dev)]
    del warm
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clk:
        e0.record()
        res = launch(steps, seed=2)
        e1.record()
        torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    evals = world * CHAINS_PER_GPU * steps * (L_STEPS + 1)
    value = evals / (ms * 1e-3)
    acc_rate = float(res.accepted.float().mean())
    # ESS/s (the metric's second half): rank-normalised split-R-hat / bulk-ESS (vihmc.diagnostics, Vehtari et al. 2021)
    # over the post-burn draws of THIS rank's chains (burn = steps // 5 as cfg.burn = num_samples // 5), per second of
    # the device-timed run; summed over ranks because chains are independent.
    ess = None
    if steps >= 40:
        from vihmc import diagnostics
        burn_rows = steps // 5
        summ = diagnostics.summarize(res.samples[burn_rows:], logp=res.logp[burn_rows:])
        t_ess = torch.tensor([summ["ess_bulk_min"], summ["ess_bulk_median"], summ["ess_bulk_logp"]], device=dev, dtype=torch.float64)
        t_rhat = torch.tensor([summ["rhat_max"]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_ess, op=dist.ReduceOp.SUM)
            dist.all_reduce(t_rhat, op=dist.ReduceOp.MAX)
        ess = {"definition": "bulk-ESS, rank-normalised, split chains; min / median over the d sampled coordinates, chains pooled",
               "post_burn_draws": int(summ["draws"]), "chains": int(summ["chains"]) * world,
               "ess_bulk_min": float(t_ess[0]), "ess_bulk_median": float(t_ess[1]), "ess_bulk_logp": float(t_ess[2]),
               "rhat_max": float(t_rhat[0]),
               "ess_min_per_sec": float(t_ess[0]) / (ms * 1e-3), "ess_median_per_sec": float(t_ess[1]) / (ms * 1e-3)}

    # ---- end to end through the public API: host tensors in, host samples out, every call ----
    # warm-up call of the SAME shape as the timed one: the pinned result buffers (53 MB at 300 iterations) then come out of
    # torch's caching host allocator instead of a fresh cudaHostAlloc (page-locking 53 MB costs ~0.1 s on a fresh box)
    samplers.sample(spec, q0_host, num_samples=steps, num_steps_per_sample=L_STEPS, step_size=STEP_SIZE, seed=3,
                    chain_offset=chain0)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = samplers.sample(spec, q0_host, num_samples=steps, num_steps_per_sample=L_STEPS, step_size=STEP_SIZE, seed=4,
                          chain_offset=chain0, return_result=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    h2d = prep.h2d_bytes + q0_host.numel() * 4
    d2h = sum(x.numel() * x.element_size() for x in (out.samples, out.accepted, out.hamiltonians, out.logp, out.step_sizes))
    e2e_value = evals / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max_mhz = peaks.get("sm_max_mhz", 1965.0)
    fp32_peak_tflops = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    achieved_tflops = (CHAINS_PER_GPU * steps * (L_STEPS + 1) * FLOP_PER_GRAD_EVAL) / (ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains_total": world * CHAINS_PER_GPU,
                   "step": "one HMC iteration of all chains = L+1 = 197 grad-evals per chain",
                   "l2": "256 MB flush before the timed launch; chain state is shared-memory resident",
                   "parallelism": f"chains sharded {world}x{CHAINS_PER_GPU}, no data-path collective"},
        "roofline": {"bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved_tflops / fp32_peak_tflops, "traffic": None,
                     "kernel": "mlp_small_sample_kernel<W=10, warps/chain=1, specialised>",
                     "note": "neither hbm nor tensor: 20x10x10 tiles are below any UMMA shape and all state lives in "
                             "shared memory; peak = 148 SM x 128 FMA lanes x 2 x sm_max_mhz (computed, not in "
                             "MEASURED_PEAKS.json); achieved = 14 kFLOP per chain-grad-eval (SURVEY 8(d)) / event time. "
                             "ncu (profiles/r01_ncu_details_mlp_small_sample_final2.csv, profiles/r01_summary.md): the binding unit is "
                             "the shared-memory pipe, l1tex lsu shared wavefronts 78 % of peak (303 per evaluation), 447 warp "
                             "instructions per evaluation, issue slots 29 % busy; DRAM 217 KB read / 0 written per 40-iteration launch"},
        # the unit that actually binds this kernel (ncu): shared-memory wavefronts.  303 wavefronts per chain-grad-eval is a constant
        # of the compiled kernel (profiles/r01_ncu_details_mlp_small_sample_final2.csv: l1tex__data_pipe_lsu_wavefronts_mem_shared
        # 2.446e9 for 1024 chains x 40 iterations x 197 evaluations); the pipe delivers one wavefront per clock per SM
        "roofline_binding_unit": {"bound": "shared-memory pipe (l1tex lsu wavefronts)", "unit": "Gwavefronts/s",
                                  "achieved": 303.0 * (CHAINS_PER_GPU * steps * (L_STEPS + 1)) / (ms * 1e-3) / 1e9,
                                  "peak": 148 * sm_max_mhz * 1e6 / 1e9,
                                  "frac": 303.0 * (CHAINS_PER_GPU * steps * (L_STEPS + 1)) / (ms * 1e-3) / (148 * sm_max_mhz * 1e6),
                                  "wavefronts_per_chain_grad_eval": 303.0},
        "cpu_baseline": cpu_baseline_leg() if (world == 1 and not skip_cpu) else None,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": d2h / steps,
                "seconds": e2e_s},
        "gpu_launches": 1,
        "clocks": clk.summary(),
        "acceptance_rate": acc_rate,
        "ess": ess,
    }
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs only")
    a = ap.parse_args()
    line = run_reference(a.steps, a.warmup, a.gpus) if a.impl == "reference" else run_ours(a.steps, a.warmup, a.gpus, a.skip_cpu_baseline)
    if line is not None:
        if _REAL_STDOUT_FD is not None:
            os.write(_REAL_STDOUT_FD, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
