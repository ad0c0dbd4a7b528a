// Bayes-by-Backprop trainer on the engine's gradient kernels -- the producer of the VI artefacts the VI-HMC path consumes.
//
// Reference: Neural_network/VI/main_regression_VI.py:75-124 (train_model), :279-346 (run), layers/BBB/BBBLinear.py:53-79 (W = mu +
// softplus(rho) * eps per weight, kl_loss), metrics.py:12-20 (ELBO = Gaussian NLL (sum) + beta * KL), :47-49 (calculate_kl);
// Operator_network/VI/main_VI_deeponet.py:56-80 (the same loop over mini-batches, NLL mean * train_size).
// One optimiser step of the reference is num_ens forward/backward passes through the Bayesian net; here it is ONE call of the
// hot-path kernel vihmc_logp_grad with C = num_ens parameter vectors W_e = mu + softplus(rho) * eps_e (a spec without prior gives
// d loglik / d W_e), bracketed by three small kernels:
//   vi_draw_kernel       W[e, i], eps[e, i] from Philox (counter = (draw e, step, i / 4, STREAM_VI_REDRAW)) or an injected stream
//   vi_step_kernel       d loss / d mu_i  = -s/E sum_e g[e, i]            + beta dKL/dmu_i
//                        d loss / d rho_i = (-s/E sum_e g[e, i] eps[e, i] + beta dKL/dsigma_i) sigmoid(rho_i)
//                        followed by torch.optim.Adam's update of (mu_i, rho_i); per-CTA partial sums of KL before / after the update
//   vi_batch_end_kernel  fixed-order sum of the partials, batch loss = -s mean_e loglik_e + beta KL, step counter
//   vi_epoch_end_kernel  validation loss at W = mu (model.eval()), torch ReduceLROnPlateau, best-checkpoint flag, history row
// Everything the host would otherwise read back (step count, learning rate, best loss) lives in a device-side state block, so an
// epoch is a fixed launch sequence without host synchronisation: the Python side captures it in a CUDA graph and replays it.
//
// KL term: BBBLinear.kl_loss calls calculate_kl(prior_mu, prior_sigma, W_mu, W_sigma) while the function is declared
// calculate_kl(mu_q, sig_q, mu_p, sig_p): the arguments are swapped, so the reference optimises KL(prior || q),
//   kl_i = 0.5 (2 log(sigma_i / sigma_p) - 1 + (sigma_p / sigma_i)^2 + ((mu_i - mu_p) / sigma_i)^2).
// kl_form = 0 reproduces that (parity with the reference's artefacts); kl_form = 1 is the textbook KL(q || prior).
#include "common.cuh"

namespace vihmc {

namespace {

constexpr int kThreads = 256;

struct VIState {
  long long step;        // optimiser steps taken (Philox counter, Adam bias correction)
  long long epoch;
  double train_acc;      // sum of batch losses of the running epoch
  int batches;           // batches accumulated in train_acc
  int num_bad;           // ReduceLROnPlateau.num_bad_epochs
  int improved;          // 1 if the epoch that just ended has the best validation loss so far
  float lr;
  float best;            // ReduceLROnPlateau.best
  float best_valid;      // checkpoint rule: valid_loss <= valid_loss_min (main_regression_VI.py:333)
  float kl_old, kl_new;  // KL before / after the last update
};

struct Layout {
  size_t state, m_mu, v_mu, m_rho, v_rho, W, eps, grad, logp, part, best_mu, best_rho, total;
  int blocks;
};

inline size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

Layout make_layout(long long D, int E) {
  Layout L{};
  L.blocks = (int)((D + kThreads - 1) / kThreads);
  size_t o = 0;
  L.state = o;    o += al(sizeof(VIState));
  L.m_mu = o;     o += al(D * sizeof(float));
  L.v_mu = o;     o += al(D * sizeof(float));
  L.m_rho = o;    o += al(D * sizeof(float));
  L.v_rho = o;    o += al(D * sizeof(float));
  L.W = o;        o += al((size_t)E * D * sizeof(float));
  L.eps = o;      o += al((size_t)E * D * sizeof(float));
  L.grad = o;     o += al((size_t)E * D * sizeof(float));
  L.logp = o;     o += al((size_t)E * sizeof(float));
  L.part = o;     o += al((size_t)2 * L.blocks * sizeof(float));
  L.best_mu = o;  o += al(D * sizeof(float));
  L.best_rho = o; o += al(D * sizeof(float));
  L.total = o;
  return L;
}

__device__ __forceinline__ float softplus(float rho) { return log1pf(expf(rho)); }   // torch.log1p(torch.exp(rho)), BBBLinear.py:57

__device__ __forceinline__ float kl_term(int form, float mu, float sg, float mu_p, float sg_p) {
  const float dm = mu - mu_p;
  if (form == 0) {
    const float r = sg_p / sg, z = dm / sg;
    return 0.5f * (2.0f * logf(sg / sg_p) - 1.0f + r * r + z * z);
  }
  const float r = sg / sg_p, z = dm / sg_p;
  return 0.5f * (2.0f * logf(sg_p / sg) - 1.0f + r * r + z * z);
}

__global__ void vi_init_kernel(VIState* s, float lr) {
  s->step = 0; s->epoch = 0; s->train_acc = 0.0; s->batches = 0; s->num_bad = 0; s->improved = 0;
  s->lr = lr; s->best = INFINITY; s->best_valid = INFINITY; s->kl_old = 0.f; s->kl_new = 0.f;
}

__global__ void __launch_bounds__(kThreads) vi_draw_kernel(const VIState* __restrict__ s, unsigned long long seed, int E, long long D,
                                                           const float* __restrict__ mu, const float* __restrict__ rho,
                                                           const float* __restrict__ inject_eps, float* __restrict__ W,
                                                           float* __restrict__ eps) {
  const long long blocks_per_row = (D + 3) / 4;
  const long long total = (long long)E * blocks_per_row;
  const long long step = s->step;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / blocks_per_row, j = t % blocks_per_row;
    float zz[4];
    if (inject_eps == nullptr) {
      const float4 z = philox_normal4(seed, (unsigned long long)e, (uint32_t)step, (uint32_t)j, STREAM_VI_REDRAW);
      zz[0] = z.x; zz[1] = z.y; zz[2] = z.z; zz[3] = z.w;
    }
    for (int k = 0; k < 4; ++k) {
      const long long i = 4 * j + k;
      if (i >= D) break;
      const float z = inject_eps != nullptr ? inject_eps[(step * E + e) * D + i] : zz[k];
      eps[e * D + i] = z;
      // weight = W_mu + W_eps * W_sigma as two rounded torch ops (BBBLinear.py:58)
      W[e * D + i] = __fadd_rn(mu[i], __fmul_rn(z, softplus(rho[i])));
    }
  }
}

__global__ void __launch_bounds__(kThreads) vi_step_kernel(VIState* __restrict__ s, int E, long long D, float nll_scale, float beta,
                                                           int kl_form, float mu_p, float sg_p, float b1, float b2, float adam_eps,
                                                           const float* __restrict__ grad, const float* __restrict__ eps,
                                                           float* __restrict__ mu, float* __restrict__ rho, float* __restrict__ m_mu,
                                                           float* __restrict__ v_mu, float* __restrict__ m_rho, float* __restrict__ v_rho,
                                                           float* __restrict__ part) {
  __shared__ float red[2][kThreads / 32];
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  float kl0 = 0.f, kl1 = 0.f;
  if (i < D) {
    float gsum = 0.f, gesum = 0.f;
    for (int e = 0; e < E; ++e) {
      const float g = grad[(long long)e * D + i];
      gsum += g;
      gesum = fmaf(g, eps[(long long)e * D + i], gesum);
    }
    const float m = mu[i], r = rho[i], sg = softplus(r), dm = m - mu_p;
    kl0 = kl_term(kl_form, m, sg, mu_p, sg_p);
    float dkl_dmu, dkl_dsg;
    if (kl_form == 0) {
      dkl_dmu = dm / (sg * sg);
      dkl_dsg = 1.0f / sg - (sg_p * sg_p + dm * dm) / (sg * sg * sg);
    } else {
      dkl_dmu = dm / (sg_p * sg_p);
      dkl_dsg = -1.0f / sg + sg / (sg_p * sg_p);
    }
    const float inv_e = 1.0f / (float)E;
    const float g_mu = -nll_scale * inv_e * gsum + beta * dkl_dmu;
    const float dsg_drho = 1.0f / (1.0f + expf(-r));   // d softplus / d rho
    const float g_rho = (-nll_scale * inv_e * gesum + beta * dkl_dsg) * dsg_drho;
    // torch.optim.Adam (amsgrad off, no weight decay)
    const float t = (float)(s->step + 1);
    const float bc1 = 1.0f - powf(b1, t), bc2s = sqrtf(1.0f - powf(b2, t));
    const float step_size = s->lr / bc1;
    float a = m_mu[i], v = v_mu[i];
    a = a + (1.0f - b1) * (g_mu - a);
    v = b2 * v + (1.0f - b2) * g_mu * g_mu;
    m_mu[i] = a; v_mu[i] = v;
    const float m_new = m - step_size * (a / (sqrtf(v) / bc2s + adam_eps));
    a = m_rho[i]; v = v_rho[i];
    a = a + (1.0f - b1) * (g_rho - a);
    v = b2 * v + (1.0f - b2) * g_rho * g_rho;
    m_rho[i] = a; v_rho[i] = v;
    const float r_new = r - step_size * (a / (sqrtf(v) / bc2s + adam_eps));
    mu[i] = m_new; rho[i] = r_new;
    kl1 = kl_term(kl_form, m_new, softplus(r_new), mu_p, sg_p);
  }
  kl0 = warp_sum(kl0); kl1 = warp_sum(kl1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = kl0; red[1][threadIdx.x >> 5] = kl1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
    part[blockIdx.x] = a;
    part[gridDim.x + blockIdx.x] = b;
  }
}

__global__ void vi_batch_end_kernel(VIState* s, int E, int blocks, float nll_scale, float beta, const float* __restrict__ logp,
                                    const float* __restrict__ part) {
  double k0 = 0.0, k1 = 0.0, ll = 0.0;
  for (int b = 0; b < blocks; ++b) { k0 += part[b]; k1 += part[blocks + b]; }
  for (int e = 0; e < E; ++e) ll += logp[e];
  s->kl_old = (float)k0; s->kl_new = (float)k1;
  s->train_acc += -(double)nll_scale * ll / E + (double)beta * k0;
  s->batches += 1;
  s->step += 1;
}

__global__ void vi_epoch_end_kernel(VIState* s, const float* __restrict__ valid_logp, int n_valid, float valid_nll_scale, float beta,
                                    int patience, float factor, float threshold, float min_lr, float* __restrict__ history) {
  double vl = 0.0;
  for (int b = 0; b < n_valid; ++b) vl += -(double)valid_nll_scale * valid_logp[b] + (double)beta * s->kl_new;
  const float valid = n_valid > 0 ? (float)(vl / n_valid) : 0.f;
  const float train = s->batches > 0 ? (float)(s->train_acc / s->batches) : 0.f;
  history[3 * s->epoch + 0] = train;
  history[3 * s->epoch + 1] = valid;
  history[3 * s->epoch + 2] = s->lr;
  // torch.optim.lr_scheduler.ReduceLROnPlateau(mode='min', threshold_mode='rel', cooldown=0, eps=1e-8).step(valid)
  if (valid < s->best * (1.0f - threshold)) { s->best = valid; s->num_bad = 0; }
  else s->num_bad += 1;
  if (s->num_bad > patience) {
    const float new_lr = fmaxf(s->lr * factor, min_lr);
    if (s->lr - new_lr > 1e-8f) s->lr = new_lr;
    s->num_bad = 0;
  }
  s->improved = valid <= s->best_valid ? 1 : 0;
  if (s->improved) s->best_valid = valid;
  s->train_acc = 0.0; s->batches = 0; s->epoch += 1;
}

__global__ void __launch_bounds__(kThreads) vi_snapshot_kernel(const VIState* __restrict__ s, long long D, const float* __restrict__ mu,
                                                               const float* __restrict__ rho, float* __restrict__ best_mu,
                                                               float* __restrict__ best_rho) {
  if (!s->improved) return;
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i < D) { best_mu[i] = mu[i]; best_rho[i] = rho[i]; }
}

}  // namespace

size_t vi_workspace_bytes(long long D, int E) { return (D < 1 || E < 1) ? 0 : make_layout(D, E).total; }

void* vi_buffer(void* ws, long long D, int E, int which) {
  const Layout L = make_layout(D, E);
  char* b = static_cast<char*>(ws);
  switch (which) {
    case 0: return b + L.W;
    case 1: return b + L.grad;
    case 2: return b + L.logp;
    case 3: return b + L.eps;
    case 4: return b + L.best_mu;
    case 5: return b + L.best_rho;
    case 6: return b + L.state;
    default: return nullptr;
  }
}

static int check_cfg(const vihmc_vi_cfg* c, long long D, void* ws, size_t ws_bytes) {
  if (c == nullptr || ws == nullptr) return fail(VIHMC_ERR_INVALID, "vi: null configuration or workspace");
  if (D < 1 || c->num_ens < 1) return fail(VIHMC_ERR_INVALID, "vi: D and num_ens must be positive");
  if (!(c->prior_sigma > 0.f)) return fail(VIHMC_ERR_INVALID, "vi: prior_sigma must be positive");
  if (ws_bytes < make_layout(D, c->num_ens).total) return fail(VIHMC_ERR_WORKSPACE, "vi: workspace too small, need %zu bytes", make_layout(D, c->num_ens).total);
  return VIHMC_OK;
}

int vi_init(const vihmc_vi_cfg* c, long long D, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (int rc = check_cfg(c, D, ws, ws_bytes)) return rc;
  const Layout L = make_layout(D, c->num_ens);
  char* b = static_cast<char*>(ws);
  VIHMC_CUDA_OK(cudaMemsetAsync(b + L.m_mu, 0, L.W - L.m_mu, st));   // the four Adam moment vectors are contiguous
  vi_init_kernel<<<1, 1, 0, st>>>(reinterpret_cast<VIState*>(b + L.state), c->lr_start);
  VIHMC_LAUNCH_OK("vi_init_kernel");
  return VIHMC_OK;
}

int vi_draw(const vihmc_vi_cfg* c, long long D, const float* mu, const float* rho, const float* inject_eps, void* ws, size_t ws_bytes,
            cudaStream_t st) {
  if (int rc = check_cfg(c, D, ws, ws_bytes)) return rc;
  if (mu == nullptr || rho == nullptr) return fail(VIHMC_ERR_INVALID, "vi_draw: null parameters");
  const Layout L = make_layout(D, c->num_ens);
  char* b = static_cast<char*>(ws);
  const long long total = (long long)c->num_ens * ((D + 3) / 4);
  const unsigned grid = (unsigned)((total + kThreads - 1) / kThreads < 148 * 8 ? (total + kThreads - 1) / kThreads : 148 * 8);
  vi_draw_kernel<<<grid, kThreads, 0, st>>>(reinterpret_cast<const VIState*>(b + L.state), c->seed, c->num_ens, D, mu, rho, inject_eps,
                                            reinterpret_cast<float*>(b + L.W), reinterpret_cast<float*>(b + L.eps));
  VIHMC_LAUNCH_OK("vi_draw_kernel");
  return VIHMC_OK;
}

int vi_step(const vihmc_vi_cfg* c, long long D, float nll_scale, float* mu, float* rho, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (int rc = check_cfg(c, D, ws, ws_bytes)) return rc;
  if (mu == nullptr || rho == nullptr) return fail(VIHMC_ERR_INVALID, "vi_step: null parameters");
  const Layout L = make_layout(D, c->num_ens);
  char* b = static_cast<char*>(ws);
  VIState* s = reinterpret_cast<VIState*>(b + L.state);
  auto f = [&](size_t off) { return reinterpret_cast<float*>(b + off); };
  vi_step_kernel<<<L.blocks, kThreads, 0, st>>>(s, c->num_ens, D, nll_scale, c->beta, c->kl_form, c->prior_mu, c->prior_sigma, c->adam_b1,
                                                c->adam_b2, c->adam_eps, f(L.grad), f(L.eps), mu, rho, f(L.m_mu), f(L.v_mu), f(L.m_rho),
                                                f(L.v_rho), f(L.part));
  VIHMC_LAUNCH_OK("vi_step_kernel");
  vi_batch_end_kernel<<<1, 1, 0, st>>>(s, c->num_ens, L.blocks, nll_scale, c->beta, f(L.logp), f(L.part));
  VIHMC_LAUNCH_OK("vi_batch_end_kernel");
  return VIHMC_OK;
}

int vi_epoch_end(const vihmc_vi_cfg* c, long long D, const float* valid_logp, int n_valid, float valid_nll_scale, const float* mu,
                 const float* rho, float* history, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (int rc = check_cfg(c, D, ws, ws_bytes)) return rc;
  if (history == nullptr || (n_valid > 0 && valid_logp == nullptr)) return fail(VIHMC_ERR_INVALID, "vi_epoch_end: null pointer");
  const Layout L = make_layout(D, c->num_ens);
  char* b = static_cast<char*>(ws);
  VIState* s = reinterpret_cast<VIState*>(b + L.state);
  vi_epoch_end_kernel<<<1, 1, 0, st>>>(s, valid_logp, n_valid, valid_nll_scale, c->beta, c->patience, c->lr_factor, c->plateau_threshold,
                                       c->min_lr, history);
  VIHMC_LAUNCH_OK("vi_epoch_end_kernel");
  vi_snapshot_kernel<<<L.blocks, kThreads, 0, st>>>(s, D, mu, rho, reinterpret_cast<float*>(b + L.best_mu),
                                                    reinterpret_cast<float*>(b + L.best_rho));
  VIHMC_LAUNCH_OK("vi_snapshot_kernel");
  return VIHMC_OK;
}

}  // namespace vihmc
