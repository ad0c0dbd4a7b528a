"""TEST INFRASTRUCTURE ONLY -- the BNN VI-HMC closure and the restated sampler for MANY chains at once, numpy fp64.

A vectorised twin of ``oracle/closures.py::BnnLogProb`` (the closure Neural_network/VI_HMC/main_VI_HMC.py:96-151 with the
functional model of my_make_func.py:52-73; 1-w-w-1 tanh, 'NLL' likelihood, isotropic 'sliced' prior) and of
``oracle/hamiltorch_restated.py::sample`` (leapfrog, H, Metropolis rule): every chain performs exactly the operations of the
per-chain oracle, but the gradient is written out by hand and evaluated for all chains in one set of numpy calls, so that
the long-run reference statistics of the posterior-parity test (64 chains x 2200 iterations x 197 evaluations) take a minute
instead of hours.  PINNED by ``tests/test_oracle_closures.py::test_batched_bnn_oracle_matches_the_pinned_closure``: values and
gradients against the torch oracle (itself pinned to the reference's closure by the golden vectors), and a whole sampling run
against ``hamiltorch_restated.sample`` with the same momenta and uniforms.
"""
from __future__ import annotations

import numpy as np


class BatchedBnn:
    def __init__(self, x, y, mu, ind, tau_out=0.0025, prior_var=1.0, width=10):
        self.x = np.asarray(x, np.float64).reshape(-1)          # [N] (in_dim = 1)
        self.y = np.asarray(y, np.float64).reshape(-1)
        self.mu = np.asarray(mu, np.float64)
        self.ind = np.asarray(ind, np.int64)
        self.v = max(float(tau_out), 1e-6)                       # GaussianNLLLoss clamps the variance at eps = 1e-6
        self.prior_var = float(prior_var)
        self.w = width
        w = width
        self.slices = {}
        off = 0
        for name, n in (("W0", w), ("b0", w), ("W1", w * w), ("b1", w), ("W2", w), ("b2", 1)):   # model.parameters() order
            self.slices[name] = slice(off, off + n)
            off += n
        self.D = off

    def _weights(self, q):
        W = np.repeat(self.mu[None], q.shape[0], 0)
        W[:, self.ind] = q                                        # my_make_func.py:56-57
        s, w = self.slices, self.w
        return W[:, s["W0"]], W[:, s["b0"]], W[:, s["W1"]].reshape(-1, w, w), W[:, s["b1"]], W[:, s["W2"]], W[:, s["b2"]][:, 0]

    def forward(self, q, x=None):
        """outputs [C, N] for parameter vectors q [C, d]"""
        x = self.x if x is None else np.asarray(x, np.float64).reshape(-1)
        W0, b0, W1, b1, W2, b2 = self._weights(q)
        h0 = np.tanh(W0[:, None, :] * x[None, :, None] + b0[:, None, :])            # [C, N, w]
        h1 = np.tanh(np.einsum("cnk,cjk->cnj", h0, W1) + b1[:, None, :])
        return np.einsum("cnj,cj->cn", h1, W2) + b2[:, None]

    def logp_grad(self, q, need_grad=True):
        """log-posterior [C] and its gradient [C, d]"""
        x, y, v = self.x, self.y, self.v
        W0, b0, W1, b1, W2, b2 = self._weights(q)
        h0 = np.tanh(W0[:, None, :] * x[None, :, None] + b0[:, None, :])
        h1 = np.tanh(np.einsum("cnk,cjk->cnj", h0, W1) + b1[:, None, :])
        o = np.einsum("cnj,cj->cn", h1, W2) + b2[:, None]
        r = o - y[None]
        ll = -0.5 * (np.log(v) * r.shape[1] + (r * r).sum(1) / v)
        lp = ll + (-0.5 * np.log(2 * np.pi * self.prior_var) * q.shape[1] - 0.5 * (q * q).sum(1) / self.prior_var)
        if not need_grad:
            return lp, None
        dO = -r / v                                                                   # [C, N]
        gW2 = np.einsum("cn,cnj->cj", dO, h1)
        gb2 = dO.sum(1, keepdims=True)
        dz1 = dO[:, :, None] * W2[:, None, :] * (1.0 - h1 * h1)
        gW1 = np.einsum("cnj,cnk->cjk", dz1, h0).reshape(q.shape[0], -1)
        gb1 = dz1.sum(1)
        dz0 = np.einsum("cnj,cjk->cnk", dz1, W1) * (1.0 - h0 * h0)
        gW0 = (dz0 * x[None, :, None]).sum(1)
        gb0 = dz0.sum(1)
        gfull = np.concatenate([gW0, gb0, gW1, gb1, gW2, gb2], 1)
        return lp, gfull[:, self.ind] - q / self.prior_var


def sample(model: BatchedBnn, q0, num_samples, num_steps, step_size, rng=None, momenta=None, uniforms=None, keep=None):
    """hamiltorch's iteration (oracle/hamiltorch_restated.py) for all chains at once; returns the final states, the acceptance
    decisions [S, C], (H0, H1) [S, C, 2] and -- when ``keep`` is given -- the states after the iterations listed in it."""
    q = np.array(q0, np.float64)
    C, d = q.shape
    acc = np.zeros((num_samples, C), bool)
    ham = np.zeros((num_samples, C, 2))
    kept = []
    keep = set() if keep is None else set(int(k) for k in keep)
    for n in range(num_samples):
        p = np.asarray(momenta[n], np.float64) if momenta is not None else rng.standard_normal((C, d))
        lp0, g = model.logp_grad(q)
        H0 = -lp0 + 0.5 * (p * p).sum(1)
        qn = q.copy()
        p = p + 0.5 * step_size * g
        for s in range(num_steps):
            qn = qn + step_size * p
            lp1, g = model.logp_grad(qn)
            p = p + step_size * g
        p = p - 0.5 * step_size * g
        H1 = -lp1 + 0.5 * (p * p).sum(1)
        u = np.asarray(uniforms[n], np.float64) if uniforms is not None else rng.random(C)
        rho = np.minimum(0.0, H0 - H1)
        a = np.isfinite(H1) & (rho >= np.log(u))
        q[a] = qn[a]
        acc[n], ham[n, :, 0], ham[n, :, 1] = a, H0, H1
        if n in keep:
            kept.append(q.copy())
    return q, acc, ham, (np.stack(kept) if kept else None)
