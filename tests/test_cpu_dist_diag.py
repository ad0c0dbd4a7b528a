"""CPU tests for the diagnostics (split-R-hat / ESS definitions) and the multi-process sharding layer (gloo, world 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vihmc import diagnostics as dg
from vihmc import dist as vd


def _ar1(S, C, d, rho, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(S, C, d, dtype=torch.float64)
    x[0] = torch.randn(C, d, generator=g, dtype=torch.float64)
    s = (1 - rho * rho) ** 0.5
    for t in range(1, S):
        x[t] = rho * x[t - 1] + s * torch.randn(C, d, generator=g, dtype=torch.float64)
    return x


def test_iid_draws_have_rhat_one_and_full_ess():
    x = _ar1(1000, 8, 5, 0.0, 0)
    r = dg.rank_split_rhat(x)
    assert (r < 1.01).all() and (r > 0.99).all()
    e = dg.bulk_ess(x)
    assert (e > 0.8 * 8000).all() and (e < 1.3 * 8000).all()


def test_ar1_ess_matches_closed_form():
    rho = 0.7
    x = _ar1(4000, 4, 6, rho, 1)
    e = dg.ess(x)
    want = 16000 * (1 - rho) / (1 + rho)
    assert ((e - want).abs() / want < 0.2).all(), (e, want)


def test_unmixed_chains_are_flagged():
    x = _ar1(500, 4, 3, 0.0, 2)
    x[:, 0] += 3.0
    assert (dg.rank_split_rhat(x) > 1.2).all()
    assert (dg.split_rhat(x) > 1.2).all()
    # a chain stuck at one value (every proposal rejected) must not produce NaN ranks
    y = _ar1(200, 3, 2, 0.0, 3)
    y[:, 1] = 0.25
    assert torch.isfinite(dg.rank_normalize(y)).all()


def test_moment_rhat_equals_draw_rhat():
    x = _ar1(600, 6, 4, 0.3, 4).float()
    m, v, S = dg.half_chain_moments(x)
    np.testing.assert_allclose(dg.rhat_from_moments(m, v, S).numpy(), dg.split_rhat(x).numpy(), rtol=1e-12)
    s = dg.summarize(x, logp=x[..., 0])
    assert set(s) >= {"rhat_max", "ess_bulk_min", "ess_bulk_median", "rhat_logp"}


def test_shard_chains_covers_everything_once():
    for total, w in ((1024, 8), (10, 3), (7, 7), (4096, 8)):
        seen = []
        for r in range(w):
            s, n = vd.shard_chains(total, r, w)
            seen += list(range(s, s + n))
        assert seen == list(range(total))
    with pytest.raises(ValueError):
        vd.shard_chains(2, 0, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = _ar1(40, total, 3, 0.2, 7).float()          # every rank can rebuild the global truth
        c0, n = vd.shard_chains(total)
        local = full[:, c0:c0 + n].contiguous()
        got = vd.gather_chains(local, total)
        rhat = vd.global_split_rhat(local)
        ok = torch.allclose(rhat, dg.split_rhat(full), rtol=1e-10)
        if rank == 0:
            ok = ok and torch.equal(got, full)
        else:
            ok = ok and got is None
        open(os.path.join(tmp, f"ok{rank}"), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_gloo_world2_gather_and_global_rhat(tmp_path, total):
    """world_size 2 on CPU: ragged shards (5 chains -> 3 + 2), gather to rank 0, R-hat from all-gathered moments."""
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / f"ok{r}").read() for r in range(2)] == ["1", "1"]
