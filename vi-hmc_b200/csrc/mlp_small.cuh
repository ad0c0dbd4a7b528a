// Small-MLP path (BNN configs): one warp per chain, everything for a chain resident in shared memory.
//
// Replaces, for the BNN configs, hamiltorch's sample()/leapfrog()/hamiltonian() loop around the
// reference closure (Neural_network/VI_HMC/main_VI_HMC.py:96-151 + my_make_func.py:52-73; call sites
// main_VI_HMC.py:379-380 and Neural_network/HMC/main_regression_hmc.py:124-127).
//
// One gradient evaluation for one chain is 14 kFLOP + 400 tanh (1-10-10-1, N=20): far below any
// UMMA tile, so this kernel runs on the FP32 pipes and is latency/issue bound, not HBM or tensor
// bound.  Design:
//   phase A  lane = data point: forward + backward through the net for that point in registers;
//            weights are read from shared memory as warp-wide broadcasts (float4 rows);
//            activations h and pre-activation gradients dz are written column-wise to shared memory.
//   phase B  lane = sampled coordinate: g_i = sum_n dz[row_i][n] * h[col_i][n] (two float4 row reads
//            per 4 data points) -- only the d sampled coordinates are ever reduced (VI-HMC subset);
//            the prior gradient and the leapfrog kick/drift are fused into the same pass, which also
//            scatters the new q_i into the full weight table (the VI-HMC masked update).
//   The whole num_samples x (L+1) loop, Philox momenta, both Hamiltonians, the Metropolis test and
//   hamiltorch's storage rule run inside ONE launch; the only HBM traffic is the stored samples.
#pragma once
#include "common.cuh"

namespace vihmc {

constexpr int kMaxHidden = 4;

struct SmallLayout {
  // weight region (floats from the chain base)
  int wbase[kMaxHidden + 1], ws[kMaxHidden + 1], bbase[kMaxHidden + 1];
  int w_total;
  // per-coordinate state, each dp floats/ints
  int q, p, g, qf, pmu, piv, meta, wpos;
  // activation region
  int act_base, xs, h, dz, dO, ones;
  int NC, dp, total;
};

struct SmallParams {
  int act, loss, last_bias, n_hidden, in_dim;
  int widths[kMaxHidden];
  long long D, d, N;
  float tau_out, inv_prior_scale, prior_sigma_scalar, prior_log_norm;
  const float *x, *y, *frozen, *prior_mu, *prior_sigma;
  const long long* sens_ind;
  SmallLayout lay;
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// W = padded hidden width the kernel is compiled for
inline SmallLayout make_layout(int W, int n_hidden, int in_dim, long long d, long long N) {
  SmallLayout L{};
  const int WSW = round_up(W, 4);
  int off = 0;
  for (int l = 0; l <= n_hidden; ++l) {
    const int rows = l < n_hidden ? W : 1;
    L.ws[l] = l == 0 ? round_up(in_dim, 4) : WSW;
    L.wbase[l] = off;
    off += rows * L.ws[l];
    L.bbase[l] = off;
    off += l < n_hidden ? WSW : 4;
  }
  L.w_total = off;
  L.dp = round_up((int)d, 4);
  L.q = off; off += L.dp;
  L.p = off; off += L.dp;
  L.g = off; off += L.dp;
  L.qf = off; off += L.dp;
  L.pmu = off; off += L.dp;
  L.piv = off; off += L.dp;
  L.meta = off; off += L.dp;
  L.wpos = off; off += L.dp;
  L.NC = N >= 32 ? 32 : round_up((int)N, 4);
  L.act_base = off;
  int a = 0;
  L.xs = a; a += in_dim * L.NC;
  L.h = a; a += n_hidden * W * L.NC;
  L.dz = a; a += n_hidden * W * L.NC;
  L.dO = a; a += L.NC;
  L.ones = a; a += L.NC;
  L.total = off + a;
  return L;
}

// full flat index f -> (position in the padded weight table, phase-B row offsets)
__device__ __forceinline__ void decode_coord(const SmallParams& P, int W, long long f, int& wpos, int& a_off, int& b_off) {
  const SmallLayout& L = P.lay;
  long long base = 0;
  wpos = 0; a_off = L.dO; b_off = L.ones;
  for (int l = 0; l <= P.n_hidden; ++l) {
    const int out_l = l < P.n_hidden ? P.widths[l] : 1;
    const int in_l = l == 0 ? P.in_dim : P.widths[l - 1];
    const int arow = l < P.n_hidden ? L.dz + l * W * L.NC : L.dO;
    const long long numel = (long long)out_l * in_l;
    if (f < base + numel) {
      const int j = (int)((f - base) / in_l), k = (int)((f - base) % in_l);
      wpos = L.wbase[l] + j * L.ws[l] + k;
      a_off = arow + (l < P.n_hidden ? j * L.NC : 0);
      b_off = l == 0 ? L.xs + k * L.NC : L.h + ((l - 1) * W + k) * L.NC;
      return;
    }
    base += numel;
    if (l < P.n_hidden || P.last_bias) {
      if (f < base + out_l) {
        const int j = (int)(f - base);
        wpos = L.bbase[l] + j;
        a_off = arow + (l < P.n_hidden ? j * L.NC : 0);
        b_off = L.ones;
        return;
      }
      base += out_l;
    }
  }
}

// One-time per-chain setup: zero the padded weight table, load frozen weights, decode coordinates,
// load the prior, load q (and scatter it into the weight table).
template <int W>
__device__ void chain_init(float* sm, const SmallParams& P, const float* q_row, int lane) {
  const SmallLayout& L = P.lay;
  for (int i = lane; i < L.act_base; i += 32) sm[i] = 0.0f;
  float* act = sm + L.act_base;
  for (int i = lane; i < L.NC; i += 32) act[L.ones + i] = 0.0f;
  __syncwarp();
  if (P.frozen != nullptr) {
    for (long long f = lane; f < P.D; f += 32) {
      int wpos, a, b;
      decode_coord(P, W, f, wpos, a, b);
      sm[wpos] = __ldg(P.frozen + f);
    }
  }
  __syncwarp();
  int* meta = reinterpret_cast<int*>(sm + L.meta);
  int* wposv = reinterpret_cast<int*>(sm + L.wpos);
  for (int i = lane; i < (int)P.d; i += 32) {
    const long long f = P.sens_ind ? __ldg(P.sens_ind + i) : (long long)i;
    int wpos, a, b;
    decode_coord(P, W, f, wpos, a, b);
    meta[i] = (a << 16) | b;
    wposv[i] = wpos;
    const float sig = P.prior_sigma ? __ldg(P.prior_sigma + i) : P.prior_sigma_scalar;
    sm[L.piv + i] = isinf(sig) ? 0.0f : 1.0f / (sig * sig);
    sm[L.pmu + i] = P.prior_mu ? __ldg(P.prior_mu + i) : 0.0f;
    const float qv = q_row[i];
    sm[L.q + i] = qv;
    sm[L.qf + i] = qv;
    sm[wpos] = qv;
  }
  __syncwarp();
}

// Phase A + phase B over all data chunks.  On return sm[g + i] holds d loglik / d q_i (no prior yet)
// and the return value is this lane's share of the log-likelihood (to be warp-summed by the caller).
template <int W, int NH>
__device__ __forceinline__ float eval_likelihood_grad(float* sm, const SmallParams& P, const Likelihood lik, int lane) {
  const SmallLayout& L = P.lay;
  constexpr int WSW = (W + 3) / 4 * 4;
  float* act = sm + L.act_base;
  const int NC = L.NC;
  float ll_lane = 0.0f;
  const int n_chunks = (int)((P.N + 31) / 32);
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const long long n = (long long)chunk * 32 + lane;
    const bool valid = n < P.N && lane < NC;
    // ---------------- phase A: lane = data point ----------------
    if (lane < NC) {
      for (int k = 0; k < P.in_dim; ++k) act[L.xs + k * NC + lane] = valid ? __ldg(P.x + n * P.in_dim + k) : 0.0f;
      act[L.ones + lane] = valid ? 1.0f : 0.0f;
      const float yv = valid ? __ldg(P.y + n) : 0.0f;
      float h[W], da[NH][W];
      // layer 0 (runtime input width)
      {
        const float* w0 = sm + L.wbase[0];
        const float* b0 = sm + L.bbase[0];
        const int ws0 = L.ws[0];
#pragma unroll
        for (int j = 0; j < W; ++j) {
          float z = b0[j];
          for (int k = 0; k < P.in_dim; ++k) z = fmaf(w0[j * ws0 + k], act[L.xs + k * NC + lane], z);
          h[j] = act_fwd(P.act, z, da[0][j]);
          act[L.h + j * NC + lane] = h[j];
        }
      }
      // hidden layers 1..NH-1
#pragma unroll
      for (int l = 1; l < NH; ++l) {
        const float* wl = sm + L.wbase[l];
        const float* bl = sm + L.bbase[l];
        float hn[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const float4* row = reinterpret_cast<const float4*>(wl + j * WSW);
          float z = bl[j];
#pragma unroll
          for (int k4 = 0; k4 < WSW / 4; ++k4) {
            const float4 w = row[k4];
            if (4 * k4 + 0 < W) z = fmaf(w.x, h[4 * k4 + 0], z);
            if (4 * k4 + 1 < W) z = fmaf(w.y, h[4 * k4 + 1], z);
            if (4 * k4 + 2 < W) z = fmaf(w.z, h[4 * k4 + 2], z);
            if (4 * k4 + 3 < W) z = fmaf(w.w, h[4 * k4 + 3], z);
          }
          hn[j] = act_fwd(P.act, z, da[l][j]);
          act[L.h + (l * W + j) * NC + lane] = hn[j];
        }
#pragma unroll
        for (int j = 0; j < W; ++j) h[j] = hn[j];
      }
      // output layer (out_dim = 1) + Gaussian likelihood
      const float* wo = sm + L.wbase[NH];
      float o = sm[L.bbase[NH]];
      float wov[W];
#pragma unroll
      for (int k4 = 0; k4 < WSW / 4; ++k4) {
        const float4 w = reinterpret_cast<const float4*>(wo)[k4];
        if (4 * k4 + 0 < W) wov[4 * k4 + 0] = w.x;
        if (4 * k4 + 1 < W) wov[4 * k4 + 1] = w.y;
        if (4 * k4 + 2 < W) wov[4 * k4 + 2] = w.z;
        if (4 * k4 + 3 < W) wov[4 * k4 + 3] = w.w;
      }
#pragma unroll
      for (int k = 0; k < W; ++k) o = fmaf(wov[k], h[k], o);
      const float r = o - yv;
      const float dO = valid ? -lik.prec * r : 0.0f;
      if (valid) ll_lane += lik.ll_const - lik.half_prec * r * r;
      act[L.dO + lane] = dO;
      // backward
      float dh[W];
#pragma unroll
      for (int k = 0; k < W; ++k) dh[k] = wov[k] * dO;
#pragma unroll
      for (int l = NH - 1; l >= 0; --l) {
        float dz[W];
#pragma unroll
        for (int j = 0; j < W; ++j) {
          dz[j] = dh[j] * da[l][j];
          act[L.dz + (l * W + j) * NC + lane] = dz[j];
        }
        if (l > 0) {
          const float* wl = sm + L.wbase[l];
#pragma unroll
          for (int k = 0; k < W; ++k) dh[k] = 0.0f;
#pragma unroll
          for (int j = 0; j < W; ++j) {
            const float4* row = reinterpret_cast<const float4*>(wl + j * WSW);
#pragma unroll
            for (int k4 = 0; k4 < WSW / 4; ++k4) {
              const float4 w = row[k4];
              if (4 * k4 + 0 < W) dh[4 * k4 + 0] = fmaf(w.x, dz[j], dh[4 * k4 + 0]);
              if (4 * k4 + 1 < W) dh[4 * k4 + 1] = fmaf(w.y, dz[j], dh[4 * k4 + 1]);
              if (4 * k4 + 2 < W) dh[4 * k4 + 2] = fmaf(w.z, dz[j], dh[4 * k4 + 2]);
              if (4 * k4 + 3 < W) dh[4 * k4 + 3] = fmaf(w.w, dz[j], dh[4 * k4 + 3]);
            }
          }
        }
      }
    }
    __syncwarp();
    // ---------------- phase B: lane = sampled coordinate ----------------
    const int* meta = reinterpret_cast<const int*>(sm + L.meta);
    for (int i = lane; i < (int)P.d; i += 32) {
      const int m = meta[i];
      const float4* A = reinterpret_cast<const float4*>(act + (m >> 16));
      const float4* B = reinterpret_cast<const float4*>(act + (m & 0xffff));
      float acc0 = 0.0f, acc1 = 0.0f;
      for (int c = 0; c < NC / 4; ++c) {
        const float4 a = A[c], b = B[c];
        acc0 = fmaf(a.x, b.x, acc0);
        acc1 = fmaf(a.y, b.y, acc1);
        acc0 = fmaf(a.z, b.z, acc0);
        acc1 = fmaf(a.w, b.w, acc1);
      }
      const float gsum = acc0 + acc1;
      sm[L.g + i] = chunk == 0 ? gsum : sm[L.g + i] + gsum;
    }
    __syncwarp();
  }
  return ll_lane;
}

// Adds the prior to sm[g] and returns this lane's share of sum_i -0.5 (q-mu)^2 / sigma^2.
__device__ __forceinline__ float add_prior(float* sm, const SmallParams& P, int lane) {
  const SmallLayout& L = P.lay;
  float lp = 0.0f;
  for (int i = lane; i < (int)P.d; i += 32) {
    const float dq = sm[L.q + i] - sm[L.pmu + i], iv = sm[L.piv + i];
    lp = fmaf(-0.5f * dq * dq, iv, lp);
    sm[L.g + i] = fmaf(-dq * iv, P.inv_prior_scale, sm[L.g + i]);
  }
  return lp;
}

// ------------------------------------------------------------------------------------------------
// kernel 1: log-posterior value + gradient for C chains (vihmc_logp_grad, MLP small path)
// ------------------------------------------------------------------------------------------------
template <int W, int NH>
__global__ void __launch_bounds__(128) mlp_small_logp_grad_kernel(SmallParams P, long long C, const float* __restrict__ q,
                                                                  float* __restrict__ logp, float* __restrict__ grad) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long chain = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (chain >= C) return;
  float* sm = smem + (size_t)warp * P.lay.total;
  chain_init<W>(sm, P, q + chain * P.d, lane);
  const Likelihood lik = make_likelihood(P.loss, P.tau_out);
  const float ll_lane = eval_likelihood_grad<W, NH>(sm, P, lik, lane);
  const float lp_lane = add_prior(sm, P, lane);
  __syncwarp();
  const float total = warp_sum(fmaf(lp_lane, P.inv_prior_scale, ll_lane)) + P.prior_log_norm * P.inv_prior_scale;
  if (lane == 0) logp[chain] = total;
  if (grad != nullptr)
    for (int i = lane; i < (int)P.d; i += 32) grad[chain * P.d + i] = sm[P.lay.g + i];
}

// ------------------------------------------------------------------------------------------------
// kernel 1b: forward only (vihmc_predict, MLP small path): out[C, N]
// ------------------------------------------------------------------------------------------------
template <int W, int NH>
__global__ void __launch_bounds__(128) mlp_small_predict_kernel(SmallParams P, long long C, const float* __restrict__ q,
                                                                float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long chain = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (chain >= C) return;
  float* sm = smem + (size_t)warp * P.lay.total;
  chain_init<W>(sm, P, q + chain * P.d, lane);
  const SmallLayout& L = P.lay;
  constexpr int WSW = (W + 3) / 4 * 4;
  for (long long n = lane; n < P.N; n += 32) {
    float h[W], dummy;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      float z = sm[L.bbase[0] + j];
      for (int k = 0; k < P.in_dim; ++k) z = fmaf(sm[L.wbase[0] + j * L.ws[0] + k], __ldg(P.x + n * P.in_dim + k), z);
      h[j] = act_fwd(P.act, z, dummy);
    }
#pragma unroll
    for (int l = 1; l < NH; ++l) {
      float hn[W];
#pragma unroll
      for (int j = 0; j < W; ++j) {
        float z = sm[L.bbase[l] + j];
#pragma unroll
        for (int k = 0; k < W; ++k) z = fmaf(sm[L.wbase[l] + j * WSW + k], h[k], z);
        hn[j] = act_fwd(P.act, z, dummy);
      }
#pragma unroll
      for (int j = 0; j < W; ++j) h[j] = hn[j];
    }
    float o = sm[L.bbase[NH]];
#pragma unroll
    for (int k = 0; k < W; ++k) o = fmaf(sm[L.wbase[NH] + k], h[k], o);
    out[chain * P.N + n] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// kernel 2: the whole HMC run for C chains in one launch (vihmc_mlp_sample)
// ------------------------------------------------------------------------------------------------
struct SampleArgs {
  vihmc_sampler_cfg cfg;
  long long C;
  const float* q0;
  float* samples;
  unsigned char* accepted;
  float* hamiltonians;
  float* logp_out;
  float* step_sizes;
  const float* inj_p;
  const float* inj_u;
};

template <int W, int NH>
__global__ void __launch_bounds__(128) mlp_small_sample_kernel(SmallParams P, SampleArgs A) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long chain = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (chain >= A.C) return;
  const SmallLayout& L = P.lay;
  float* sm = smem + (size_t)warp * L.total;
  const int d = (int)P.d;
  const long long C = A.C;
  const float* q0 = A.q0 + chain * d;
  chain_init<W>(sm, P, q0, lane);
  const int* wposv = reinterpret_cast<const int*>(sm + L.wpos);
  const Likelihood lik = make_likelihood(P.loss, P.tau_out);
  const unsigned long long gchain = (unsigned long long)(A.cfg.chain_offset + chain);
  const int S = A.cfg.num_samples, nsteps = A.cfg.num_steps, burn = A.cfg.burn;
  const float log_norm = P.prior_log_norm * P.inv_prior_scale;

  // dual-averaging state (Sampler.HMC_NUTS): every lane carries the same scalars
  float eps = A.cfg.step_size;
  const float eps_init = A.cfg.step_size;
  float eps_bar = 1.0f, H_t = 0.0f;

  // stored row 0 = params_init
  for (int i = lane; i < d; i += 32) A.samples[chain * d + i] = sm[L.q + i];
  float logp_init = 0.0f, logp_f = 0.0f;  // log-posterior of params_init / of the fallback state

  for (int n = 0; n < S; ++n) {
    if (A.cfg.hamiltorch_fallback_rule && n == burn + 1) {
      for (int i = lane; i < d; i += 32) sm[L.qf + i] = q0[i];
      logp_f = logp_init;
    }
    // ---- momentum ----
    float ke = 0.0f;
    if (A.inj_p != nullptr) {
      const float* src = A.inj_p + ((long long)n * C + chain) * d;
      for (int i = lane; i < d; i += 32) {
        const float pv = src[i];
        sm[L.p + i] = pv;
        ke = fmaf(pv, pv, ke);
      }
    } else {
      for (int j = lane; 4 * j < d; j += 32) {
        const float4 z = philox_normal4(A.cfg.seed, gchain, (uint32_t)n, (uint32_t)j, STREAM_MOMENTUM);
        const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (4 * j + t < d) {
            sm[L.p + 4 * j + t] = zz[t];
            ke = fmaf(zz[t], zz[t], ke);
          }
      }
    }
    __syncwarp();
    // ---- H0 and first half kick + first drift ----
    float ll_lane = eval_likelihood_grad<W, NH>(sm, P, lik, lane);
    float lp_lane = add_prior(sm, P, lane);
    __syncwarp();
    const float logp0 = warp_sum(fmaf(lp_lane, P.inv_prior_scale, ll_lane)) + log_norm;
    const float H0 = -logp0 + 0.5f * warp_sum(ke);
    if (n == 0) {
      logp_init = logp0;
      logp_f = logp0;
      if (A.logp_out != nullptr && lane == 0) A.logp_out[chain] = logp0;
    }
    const float half_eps = 0.5f * eps;
    for (int i = lane; i < d; i += 32) {
      const float pv = axpy_unfused(half_eps, sm[L.g + i], sm[L.p + i]);
      const float qv = axpy_unfused(eps, pv, sm[L.q + i]);
      sm[L.p + i] = pv;
      sm[L.q + i] = qv;
      sm[wposv[i]] = qv;
    }
    __syncwarp();
    // ---- L leapfrog steps ----
    float logp1 = 0.0f, ke1 = 0.0f;
    for (int s = 1; s <= nsteps; ++s) {
      ll_lane = eval_likelihood_grad<W, NH>(sm, P, lik, lane);
      if (s < nsteps) {
        for (int i = lane; i < d; i += 32) {
          const float dq = sm[L.q + i] - sm[L.pmu + i];
          const float gi = fmaf(-dq * sm[L.piv + i], P.inv_prior_scale, sm[L.g + i]);
          const float pv = axpy_unfused(eps, gi, sm[L.p + i]);
          const float qv = axpy_unfused(eps, pv, sm[L.q + i]);
          sm[L.p + i] = pv;
          sm[L.q + i] = qv;
          sm[wposv[i]] = qv;
        }
      } else {
        lp_lane = 0.0f;
        for (int i = lane; i < d; i += 32) {
          const float dq = sm[L.q + i] - sm[L.pmu + i], iv = sm[L.piv + i];
          lp_lane = fmaf(-0.5f * dq * dq, iv, lp_lane);
          const float gi = fmaf(-dq * iv, P.inv_prior_scale, sm[L.g + i]);
          float pv = axpy_unfused(eps, gi, sm[L.p + i]);
          pv = __fsub_rn(pv, __fmul_rn(half_eps, gi));
          sm[L.p + i] = pv;
          ke1 = fmaf(pv, pv, ke1);
        }
        logp1 = warp_sum(fmaf(lp_lane, P.inv_prior_scale, ll_lane)) + log_norm;
        ke1 = 0.5f * warp_sum(ke1);
      }
      __syncwarp();
    }
    if (nsteps == 0) {  // degenerate: proposal == current state, p = p + eps/2 g - eps/2 g
      logp1 = logp0;
      for (int i = lane; i < d; i += 32) ke1 = fmaf(sm[L.p + i], sm[L.p + i], ke1);
      ke1 = 0.5f * warp_sum(ke1);
    }
    const float H1 = -logp1 + ke1;
    // ---- Metropolis test (hamiltorch: rho = min(0, H0-H1); accept iff rho >= log u) ----
    const float u = A.inj_u != nullptr ? A.inj_u[(long long)n * C + chain] : philox_uniform(A.cfg.seed, gchain, (uint32_t)n);
    const float rho = fminf(0.0f, H0 - H1);
    const bool finite = isfinite(logp1) && isfinite(H1) && isfinite(H0);
    const bool accept = finite && (rho >= logf(u));
    const bool store = n > burn;
    float* row = store ? A.samples + ((long long)(n - burn) * C + chain) * d : nullptr;
    if (accept) {
      logp_f = logp1;
      for (int i = lane; i < d; i += 32) {
        const float qv = sm[L.q + i];
        sm[L.qf + i] = qv;
        if (store) row[i] = qv;
      }
    } else {
      for (int i = lane; i < d; i += 32) {
        const float qv = sm[L.qf + i];
        sm[L.q + i] = qv;
        sm[wposv[i]] = qv;
        if (store) row[i] = qv;
      }
    }
    if (lane == 0) {
      if (A.accepted != nullptr) A.accepted[(long long)n * C + chain] = accept ? 1 : 0;
      if (A.hamiltonians != nullptr) {
        A.hamiltonians[((long long)n * C + chain) * 2 + 0] = H0;
        A.hamiltonians[((long long)n * C + chain) * 2 + 1] = H1;
      }
      if (A.logp_out != nullptr && store) A.logp_out[(long long)(n - burn) * C + chain] = logp_f;
    }
    // ---- dual averaging (hamiltorch adaptation(): gamma .05, t0 10, kappa .75, mu = log(10 eps0)) ----
    if (A.cfg.adapt_step_size && n <= burn) {
      if (n < burn) {
        const float t = (float)(n + 1);
        const float alpha = finite ? fminf(1.0f, expf(rho)) : 0.0f;
        const float mu = logf(10.0f * eps_init);
        H_t = (1.0f - 1.0f / (t + 10.0f)) * H_t + (1.0f / (t + 10.0f)) * (A.cfg.desired_accept_rate - alpha);
        const float x_new = mu - sqrtf(t) / 0.05f * H_t;
        eps = expf(x_new);
        const float tk = powf(t, -0.75f);
        eps_bar = expf(tk * x_new + (1.0f - tk) * logf(eps_bar));
      }
      if (n == burn) eps = eps_bar;
    }
    __syncwarp();
  }
  if (A.step_sizes != nullptr && lane == 0) A.step_sizes[chain] = eps;
}


enum SmallOp { kOpLogpGrad = 0, kOpPredict = 1, kOpSample = 2 };

struct SmallLaunch {
  int warps_per_block, blocks;
  size_t smem;
  long long C;
  const float* q;
  float* logp;
  float* grad;
  float* out;
  SampleArgs A;
};

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48u * 1024u) VIHMC_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return VIHMC_OK;
}

template <int W, int NH>
static int launch_small_wn(SmallOp op, const SmallParams& P, const SmallLaunch& a, cudaStream_t st) {
  const int threads = a.warps_per_block * 32;
  if (op == kOpLogpGrad) {
    auto k = mlp_small_logp_grad_kernel<W, NH>;
    if (int rc = set_smem(k, a.smem)) return rc;
    k<<<a.blocks, threads, a.smem, st>>>(P, a.C, a.q, a.logp, a.grad);
    VIHMC_LAUNCH_OK("mlp_small_logp_grad_kernel");
  } else if (op == kOpPredict) {
    auto k = mlp_small_predict_kernel<W, NH>;
    if (int rc = set_smem(k, a.smem)) return rc;
    k<<<a.blocks, threads, a.smem, st>>>(P, a.C, a.q, a.out);
    VIHMC_LAUNCH_OK("mlp_small_predict_kernel");
  } else {
    auto k = mlp_small_sample_kernel<W, NH>;
    if (int rc = set_smem(k, a.smem)) return rc;
    k<<<a.blocks, threads, a.smem, st>>>(P, a.A);
    VIHMC_LAUNCH_OK("mlp_small_sample_kernel");
  }
  return VIHMC_OK;
}

template <int W>
static int launch_small_w(SmallOp op, const SmallParams& P, const SmallLaunch& a, cudaStream_t st) {
  switch (P.n_hidden) {
    case 1: return launch_small_wn<W, 1>(op, P, a, st);
    case 2: return launch_small_wn<W, 2>(op, P, a, st);
    case 3: return launch_small_wn<W, 3>(op, P, a, st);
    case 4: return launch_small_wn<W, 4>(op, P, a, st);
  }
  return fail(VIHMC_ERR_UNSUPPORTED, "no small-MLP instantiation for %d hidden layers", P.n_hidden);
}

}  // namespace vihmc
