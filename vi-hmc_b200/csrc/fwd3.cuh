// Forward pass of the dense DeepONet path on the exact-accumulation products of xgemm.cuh:
//   weight_image3_kernel     per chain and layer: W_l -> three-piece bf16 image + row scales + bias (one blob per layer)
//   fused_forward3_kernel    all layers of one stack for a 128-row tile in one kernel (the successor of
//                            fused_stack.cuh's fused_forward_kernel, same structure); the activations of the tile stay in
//                            shared memory AS the three-piece operand image of the next layer, and the last layer's image
//                            goes to HBM as the head's operand
//   head3_kernel             out[n, p] = <xb[n], xtr[p]> + b0 with the Gaussian residual epilogue: persistent, warp-specialised
//                            (one bulk-copy producer thread, one MMA-issuing thread, 8 epilogue warps), A tile resident, B tiles
//                            double-buffered in shared memory, accumulators double-buffered in tensor memory
// Restates Operator_network/VI_HMC/my_make_func.py:53-82 per chain (branch :53-61, trunk :69-77, head einsum :79, bias :81-82)
// and the likelihood of main_VI_HMC_burgers.py:157-163.
#pragma once
#include "fused_stack.cuh"
#include "xgemm.cuh"

namespace vihmc {
namespace xg {

// ---------------------------------------------------------------------------------------------------------------
// tanh with an UNBIASED error.  The forward pass's other source of coherent error: a tanh that is off by a consistent 3e-8
// relative moves the BASELINE-size gradient by 4.4e-5, so the mean error has to stay below ~1e-9 (measured by emulating the
// candidates in numpy, profiles/r02_summary.md section A: tanhf-style 1 - 2/(exp(2x)+1) on MUFU.EX2: -6e-9; this function:
// < 6e-11 for pre-activations of standard deviation 0.03 ... 2, maximum error 2.4 ulp).
//   tanh|x| = em1 / (em1 + 2), em1 = expm1(2|x|) = 2^n (1 + q) - 1:
//   n = rint(2|x| log2 e) by the magic-number add; r = 2|x| - n ln2 by Cody-Waite with a 12-bit ln2_hi (first step exact; the
//   second one rounds, and because r1 sits on the grid of 2|x| that rounding error is the SAME for a whole binade -- a bias of
//   0.07 ulp -- so its exact residual rl is carried along); q = expm1(r) = r + r^2 P(r), P the degree-6 Taylor polynomial of
//   (e^r - 1 - r) / r^2 (truncation r^9/9! <= 2e-10 relative); the quotient by one Newton step on rcp.approx and a residual
//   correction into which rl enters BEFORE the final rounding (a correction added to an already rounded fp32 value is lost).
// Every step is a round-to-nearest fp32 operation; evaluated on pairs with the packed FP32 instructions (fma.rn.f32x2).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void tanh_acc2(float x0, float x1, float& y0, float& y1) {
  using tc::fma2;
  using tc::pk2;
  const float a0 = fminf(fabsf(x0), 12.0f), a1 = fminf(fabsf(x1), 12.0f);   // tanh(12) rounds to 1
  const unsigned long long one = pk2(1.0f, 1.0f), mone = pk2(-1.0f, -1.0f);
  const unsigned long long ax2 = pk2(a0 + a0, a1 + a1);
  const unsigned long long t = fma2(ax2, pk2(1.4426950408889634f, 1.4426950408889634f), pk2(12582912.0f, 12582912.0f));
  const unsigned long long n = add2(t, pk2(-12582912.0f, -12582912.0f));
  const unsigned long long NHI = pk2(-0.693145751953125f, -0.693145751953125f), NLO = pk2(-1.42860677e-6f, -1.42860677e-6f);
  const unsigned long long r1 = fma2(n, NHI, ax2);                     // exact
  const unsigned long long r = fma2(n, NLO, r1);
  const unsigned long long rl = fma2(n, NLO, fma2(r, mone, r1));       // (r1 - r) - n ln2_lo: what the rounding of r dropped
  unsigned long long p = pk2(2.48015873015873e-05f, 2.48015873015873e-05f);               // 1/8!
  p = fma2(p, r, pk2(1.984126984126984e-04f, 1.984126984126984e-04f));                  // 1/7!
  p = fma2(p, r, pk2(1.388888888888889e-03f, 1.388888888888889e-03f));                  // 1/6!
  p = fma2(p, r, pk2(8.333333333333333e-03f, 8.333333333333333e-03f));                  // 1/5!
  p = fma2(p, r, pk2(4.166666666666666e-02f, 4.166666666666666e-02f));                  // 1/4!
  p = fma2(p, r, pk2(1.666666666666667e-01f, 1.666666666666667e-01f));                  // 1/3!
  p = fma2(p, r, pk2(0.5f, 0.5f));
  const unsigned long long q = fma2(mul2(r, r), p, r);
  float t0, t1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
  const unsigned long long two_n = pk2(__int_as_float((__float_as_int(t0) - 0x4B400000 + 127) << 23),
                                       __int_as_float((__float_as_int(t1) - 0x4B400000 + 127) << 23));
  const unsigned long long em1 = fma2(q, two_n, add2(two_n, mone));
  const unsigned long long elo = mul2(add2(em1, one), rl);
  const unsigned long long d = add2(em1, pk2(2.0f, 2.0f));
  float d0, d1, c0, c1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(c1) : "f"(d1));
  unsigned long long rc = pk2(c0, c1);
  const unsigned long long nd = mul2(d, mone);
  rc = fma2(rc, fma2(nd, rc, one), rc);                                 // Newton: rc += rc (1 - d rc)
  unsigned long long y = mul2(em1, rc);
  unsigned long long res = fma2(y, nd, em1);                            // em1 - y d
  res = fma2(elo, fma2(y, mone, one), res);                             // + elo (1 - y): numerator and denominator both carry elo
  y = fma2(rc, res, y);
  float b0, b1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(b0), "=f"(b1) : "l"(y));
  y0 = copysignf(b0, x0);
  y1 = copysignf(b1, x1);
}

template <int ACT>
__device__ __forceinline__ void activate8(float (&v)[8]) {
  if (ACT == VIHMC_ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 8; j += 2) tanh_acc2(v[j], v[j + 1], v[j], v[j + 1]);
  } else if (ACT == VIHMC_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.0f ? v[j] : 0.0f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight images: blob of layer l of chain c = [4 pieces x XPIECE_B | 112 row scales | 112 bias values | pad]
// FOUR pieces (31 bits below the row scale): the weights are INPUTS -- exact fp32 numbers that the reference uses as they are,
// and the same for every data row.  Rounding them to 24 bits below the row's largest entry is a perturbation of the evaluation
// point that is coherent across the whole data set: measured (CPU emulation of this scheme, profiles/r02_summary.md section A)
// it alone shifts the BASELINE-size outputs by -3.0e-7 coherently (5e-5 on the gradient), whereas the same rounding of the
// activations, which differs from row to row, shifts them by 9e-9.  With four pieces every weight above 2^-8 of its row's
// maximum is represented exactly; the products with i + j <= 5 go to FOUR exact accumulators (L3 = p1 q4 + p2 q3 + p3 q2).
// ---------------------------------------------------------------------------------------------------------------
constexpr int NPADX = 112;                          // accumulator columns (output units, padded)
constexpr int XPIECE_B = (NPADX / 8) * XRG;         // 25 088 B: one piece of a 112-row operand
constexpr int XSC_OFF = 4 * XPIECE_B;               // row scales
constexpr int XBIAS_OFF = XSC_OFF + NPADX * 4;      // bias
constexpr int XIMG_B = (XBIAS_OFF + NPADX * 4 + 127) / 128 * 128;   // 101 248 B -> 101 376

struct Img3Layer { long long w_off, b_off; int n_rows, n_k, ld_n, ld_k; };   // element (n, k) = W[w_off + n ld_n + k ld_k]; b_off < 0: no bias
struct Img3Table { int n; Img3Layer L[VIHMC_MAX_LAYERS]; };

__global__ void __launch_bounds__(256) weight_image3_kernel(const float* __restrict__ Wf, long long Dp, Img3Table t,
                                                            unsigned char* __restrict__ img) {
  const int l = blockIdx.x;
  const long long c = blockIdx.y;
  const Img3Layer L = t.L[l];
  const float* __restrict__ W = Wf + c * Dp + L.w_off;
  unsigned char* out = img + (c * t.n + l) * (long long)XIMG_B;
  const int row = threadIdx.x >> 1, half = threadIdx.x & 1;
  if (row >= NPADX) return;   // whole warps: 224 active threads
  const bool rv = row < L.n_rows;
  const float* __restrict__ wr = W + (long long)row * L.ld_n;
  float m = 0.0f;
  if (rv)
    for (int k = half * (XK / 2); k < (half + 1) * (XK / 2) && k < L.n_k; ++k) m = fmaxf(m, fabsf(__ldg(wr + (long long)k * L.ld_k)));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  const float s = pow2_ceil(m), inv = pow2_inv(s);
  if (half == 0) {
    reinterpret_cast<float*>(out + XSC_OFF)[row] = s;
    reinterpret_cast<float*>(out + XBIAS_OFF)[row] = (rv && L.b_off >= 0) ? __ldg(Wf + c * Dp + L.b_off + row) : 0.0f;
  }
  for (int ch = half * (XCH / 2); ch < (half + 1) * (XCH / 2); ++ch) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = ch * 8 + j;
      x[j] = (rv && k < L.n_k) ? __ldg(wr + (long long)k * L.ld_k) * inv : 0.0f;
    }
    uint32_t w[4][4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float p[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        float r;
        split3(x[j + u], p[u][0], p[u][1], p[u][2]);
        r = __fadd_rn(__fadd_rn(__fadd_rn(x[j + u], -p[u][0]), -p[u][1]), -p[u][2]);        // exact, |r| <= 2^-24
        p[u][3] = __fadd_rn(__fadd_rn(r, 0.005859375f), -0.005859375f);                     // 1.5 * 2^-8: multiple of 2^-31
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i][j >> 1] = bf16x2(p[0][i], p[1][i]);
    }
    const int off = piece_off(row, ch);
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(out + i * XPIECE_B + off) = make_uint4(w[i][0], w[i][1], w[i][2], w[i][3]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// fused stack forward
// ---------------------------------------------------------------------------------------------------------------
constexpr int F3_THREADS = 512;
constexpr int F3_SA_OFF = XTILE + XIMG_B;            // float[128]: row scales of the operand tile
constexpr int F3_RMAX_OFF = F3_SA_OFF + 512;         // float[4][128]: row maxima of the four column quarters
constexpr int F3_BAR_OFF = F3_RMAX_OFF + 2048;
constexpr int F3_SMEM = F3_BAR_OFF + 128;
constexpr int MAX_IN0 = 8;                           // widest first-layer input handled on the FP32 pipes

struct Fwd3Args {
  const float* input;                      // [R, in_dim], shared by every chain
  int in_dim;
  const float* Wf;                         // padded weights [Cb, Dp]: the narrow first layer reads W_0 and b_0 from here
  long long Dp, w0_off, b0_off;
  int ldw0;
  int dims[VIHMC_MAX_LAYERS];
  int n_layers;
  const unsigned char* wimg;               // blobs of layers l0 .. n_layers-1 (l0 = 0 for a wide input, else 1): [Cb, n, XIMG_B]
  float* acts[VIHMC_MAX_LAYERS];           // acts[l]: [Cb, R, dims[l]] fp32 (the backward pass reads them)
  long long R;
  unsigned char* out_img;                  // [Cb, tiles, XTILE]: operand image of the last layer's output (null: not needed)
  float* out_scales;                       // [Cb, tiles * 128]
};

template <int ACT>
__global__ void __launch_bounds__(F3_THREADS, 1) fused_forward3_kernel(Fwd3Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* At = smem;                              // three pieces, XPIECE each
  unsigned char* Bt = smem + XTILE;                      // the layer's blob
  float* sa = reinterpret_cast<float*>(smem + F3_SA_OFF);
  float* rmax = reinterpret_cast<float*>(smem + F3_RMAX_OFF);
  uint64_t* bar_b = reinterpret_cast<uint64_t*>(smem + F3_BAR_OFF);
  uint64_t* bar_mma = bar_b + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long c = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * 128;
  const float* __restrict__ Wc = a.Wf + c * a.Dp;
  const int l0 = a.in_dim > MAX_IN0 ? 0 : 1;
  const int n_img = a.n_layers - l0;

  if (tid == 0) {
    tc::mbar_init(bar_b, 1);
    tc::mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < XTILE / 16; i += F3_THREADS) reinterpret_cast<uint4*>(At)[i] = make_uint4(0u, 0u, 0u, 0u);   // K padding stays zero
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  auto load_weights = [&](int l) {   // thread 0: one bulk copy of layer l's blob, completion on bar_b
    expect_tx(bar_b, (uint32_t)XIMG_B);
    bulk_load(Bt, a.wimg + (c * n_img + (l - l0)) * (long long)XIMG_B, (uint32_t)XIMG_B, bar_b);
  };
  if (tid == 0 && l0 < a.n_layers) load_weights(l0);

  {  // ---- first operand: warp -> row group, lane -> (row r8 = lane % 8, chunks lane / 8 + 4 i) ----
    const int r8 = lane & 7, cq = lane >> 3, rg = warp, row = rg * 8 + r8;
    const long long grow = r0 + row;
    const bool rvalid = grow < a.R;
    constexpr int NCH = (XCH + 3) / 4;
    float v[NCH][8];
    const int width = l0 == 0 ? a.in_dim : a.dims[0];
    if (l0 == 0) {   // wide input (the branch's sensors): the input tile itself is the first operand
      const float* __restrict__ src = a.input + grow * a.in_dim;
#pragma unroll
      for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = (cq + 4 * i) * 8 + j;
          v[i][j] = (rvalid && k < a.in_dim) ? __ldg(src + k) : 0.0f;
        }
    } else {         // narrow input (the trunk's 5 features): layer 0 on the FP32 pipes
      float f[MAX_IN0];
#pragma unroll
      for (int k = 0; k < MAX_IN0; ++k) f[k] = (k < a.in_dim && rvalid) ? __ldg(a.input + grow * a.in_dim + k) : 0.0f;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int ch = cq + 4 * i;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = ch * 8 + j;
          float acc = 0.0f;
          if (n < width) {
            const float4* wrow = reinterpret_cast<const float4*>(Wc + a.w0_off + (long long)n * a.ldw0);   // rows padded to 4 floats
            float w[MAX_IN0];
#pragma unroll
            for (int k4 = 0; k4 < MAX_IN0 / 4; ++k4) {
              const float4 t = 4 * k4 < a.ldw0 ? __ldg(wrow + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
              w[4 * k4] = t.x; w[4 * k4 + 1] = t.y; w[4 * k4 + 2] = t.z; w[4 * k4 + 3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < MAX_IN0; ++k)
              if (k < a.in_dim) acc = fmaf(w[k], f[k], acc);
            acc += __ldg(Wc + a.b0_off + n);
          }
          v[i][j] = acc;
        }
        if (a.n_layers > 1) activate8<ACT>(v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (ch * 8 + j >= width || !rvalid) v[i][j] = 0.0f;
      }
    }
    float s = 1.0f;
    if (l0 == 0 || ACT != VIHMC_ACT_TANH || a.n_layers == 1) {   // tanh outputs are bounded by 1: scale 1
      float m = 0.0f;
#pragma unroll
      for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(v[i][j]));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
      s = pow2_ceil(m);
    }
    if (cq == 0) sa[row] = s;
    const float inv = pow2_inv(s);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = cq + 4 * i;
      if (ch >= XCH || ch * 8 >= width) continue;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = v[i][j] * inv;
      uint4 c1, c2, c3;
      split3_chunk(x, c1, c2, c3);
      const int off = piece_off(row, ch);
      *reinterpret_cast<uint4*>(At + off) = c1;
      *reinterpret_cast<uint4*>(At + XPIECE + off) = c2;
      *reinterpret_cast<uint4*>(At + 2 * XPIECE + off) = c3;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  // coalesced write of the activations held in the operand tile: acts[l][c, r0 + row, :] = s (p1 + p2 + p3), exact
  auto store_acts = [&](int l) {
    const int r8 = lane & 7, cq = lane >> 3, rg = warp, row = rg * 8 + r8;
    const int nw = a.dims[l];
    const long long grow = r0 + row;
    if (grow >= a.R) return;
    const float s = sa[row];
    float* __restrict__ out = a.acts[l] + (c * a.R + grow) * nw;
    for (int ch = cq; ch * 8 < nw; ch += 4) {
      const int off = piece_off(row, ch);
      const uint4 c1 = *reinterpret_cast<const uint4*>(At + off), c2 = *reinterpret_cast<const uint4*>(At + XPIECE + off),
                  c3 = *reinterpret_cast<const uint4*>(At + 2 * XPIECE + off);
      const uint32_t w1[4] = {c1.x, c1.y, c1.z, c1.w}, w2[4] = {c2.x, c2.y, c2.z, c2.w}, w3[4] = {c3.x, c3.y, c3.z, c3.w};
      float h[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        h[2 * j] = ((__uint_as_float(w1[j] << 16) + __uint_as_float(w2[j] << 16)) + __uint_as_float(w3[j] << 16)) * s;
        h[2 * j + 1] = ((__uint_as_float(w1[j] & 0xFFFF0000u) + __uint_as_float(w2[j] & 0xFFFF0000u)) + __uint_as_float(w3[j] & 0xFFFF0000u)) * s;
      }
      *reinterpret_cast<float4*>(out + ch * 8) = make_float4(h[0], h[1], h[2], h[3]);
      if (ch * 8 + 4 < nw) *reinterpret_cast<float4*>(out + ch * 8 + 4) = make_float4(h[4], h[5], h[6], h[7]);
    }
  };

  for (int l = l0; l < a.n_layers; ++l) {
    const int K = l == 0 ? a.in_dim : a.dims[l - 1], N = a.dims[l];
    const uint32_t ph = (uint32_t)(l - l0) & 1u;
    const bool last = l == a.n_layers - 1;
    if (tid == 0) {
      tc::mbar_wait(bar_b, ph);   // this layer's blob has landed
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int ksteps = (K + 15) / 16;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t ko = (uint32_t)ks * 256u, acc = ks == 0 ? 0u : 1u, id = idesc_bf16(NPADX);
        uint64_t da[3], db[4];
#pragma unroll
        for (int i = 0; i < 3; ++i) da[i] = tc::make_desc(tc::smem_u32(At) + i * XPIECE + ko, 128, XRG);
#pragma unroll
        for (int i = 0; i < 4; ++i) db[i] = tc::make_desc(tc::smem_u32(Bt) + i * XPIECE_B + ko, 128, XRG);
        mma_bf16(tmem_d, da[0], db[0], id, acc);
        mma_bf16(tmem_d + 128u, da[0], db[1], id, acc);
        mma_bf16(tmem_d + 128u, da[1], db[0], id, 1u);
        mma_bf16(tmem_d + 256u, da[0], db[2], id, acc);
        mma_bf16(tmem_d + 256u, da[1], db[1], id, 1u);
        mma_bf16(tmem_d + 256u, da[2], db[0], id, 1u);
        mma_bf16(tmem_d + 384u, da[0], db[3], id, acc);
        mma_bf16(tmem_d + 384u, da[1], db[2], id, 1u);
        mma_bf16(tmem_d + 384u, da[2], db[1], id, 1u);
      }
      tc::mma_commit(bar_mma);
    }
    if (l >= 1) store_acts(l - 1);   // reads the operand tile while the tensor core reads it too
    tc::mbar_wait(bar_b, ph);        // scales and bias of this layer are in the blob (every thread reads them below)
    tc::mbar_wait(bar_mma, ph);      // accumulators complete; operand tile is free
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, cq4 = warp >> 2;          // TMEM lane quarter; 32-column quarter of the accumulator
    const int row = q * 32 + lane;
    const int nb = cq4 * 32;
    const float sa_row = sa[row];
    float v[4][8];
    {  // epilogue part 1: TMEM (lane = row) -> z = s_a s_w (L0 + (L1 + (L2 + L3))) + b -> activation
      const float* __restrict__ sw = reinterpret_cast<const float*>(Bt + XSC_OFF);
      const float* __restrict__ bias = reinterpret_cast<const float*>(Bt + XBIAS_OFF);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t l0r[2][8], l1r[2][8], l2r[2][8], l3r[2][8];
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const int n = nb + hh * 16 + 8 * g2;
          if (n < N) {
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)n;
            tmem_ld8(taddr, l0r[g2]);
            tmem_ld8(taddr + 128u, l1r[g2]);
            tmem_ld8(taddr + 256u, l2r[g2]);
            tmem_ld8(taddr + 384u, l3r[g2]);
          }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const int g = hh * 2 + g2, n = nb + 8 * g;
          if (n < N) {
            const float4 s0 = *reinterpret_cast<const float4*>(sw + n), s1 = *reinterpret_cast<const float4*>(sw + n + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(bias + n), b1 = *reinterpret_cast<const float4*>(bias + n + 4);
            const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float low = __fadd_rn(__uint_as_float(l2r[g2][j]), __uint_as_float(l3r[g2][j]));
              const float sum = __fadd_rn(__uint_as_float(l0r[g2][j]), __fadd_rn(__uint_as_float(l1r[g2][j]), low));
              v[g][j] = fmaf(sum, sa_row * ss[j], bb[j]);
            }
            if (!last) activate8<ACT>(v[g]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[g][j] = 0.0f;
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();              // every thread has read sa, the blob and (store_acts) the operand tile
    if (tid == 0 && l + 1 < a.n_layers) load_weights(l + 1);
    // epilogue part 2: row scale of the new operand (tanh outputs: 1), pieces into the operand tile of the next layer
    float s = 1.0f;
    if (last || ACT != VIHMC_ACT_TANH) {
      float m = 0.0f;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int j = 0; j < 8; ++j) m = fmaxf(m, fabsf(v[g][j]));
      rmax[cq4 * 128 + row] = m;
      __syncthreads();
      s = pow2_ceil(fmaxf(fmaxf(rmax[row], rmax[128 + row]), fmaxf(rmax[256 + row], rmax[384 + row])));
    }
    if (cq4 == 0) sa[row] = s;
    const float inv = pow2_inv(s);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = nb + 8 * g;
      if (n < N) {
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = v[g][j] * inv;
        uint4 c1, c2, c3;
        split3_chunk(x, c1, c2, c3);
        const int off = piece_off(row, n >> 3);
        *reinterpret_cast<uint4*>(At + off) = c1;
        *reinterpret_cast<uint4*>(At + XPIECE + off) = c2;
        *reinterpret_cast<uint4*>(At + 2 * XPIECE + off) = c3;
      } else if (n < K) {   // a narrower layer leaves stale columns of the previous one: clear them (never taken at equal widths)
        const int off = piece_off(row, n >> 3);
        *reinterpret_cast<uint4*>(At + off) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(At + XPIECE + off) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(At + 2 * XPIECE + off) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();
  }
  store_acts(a.n_layers - 1);
  if (a.out_img != nullptr) {   // the tile's image and row scales are the head's operand
    uint4* dst = reinterpret_cast<uint4*>(a.out_img + (c * gridDim.x + blockIdx.x) * (long long)XTILE);
    for (int i = tid; i < XTILE / 16; i += F3_THREADS) dst[i] = reinterpret_cast<const uint4*>(At)[i];
    if (tid < 128) a.out_scales[(c * gridDim.x + blockIdx.x) * 128 + tid] = sa[tid];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// head: persistent, warp-specialised
// ---------------------------------------------------------------------------------------------------------------
constexpr int H3_THREADS = 576;                      // warp 0 producer, warp 1 MMA issuer, warps 2..17 epilogue
constexpr int H3_BN = 64;                            // trunk points per tile
constexpr int H3_BPIECE = (H3_BN / 8) * XRG;         // 14 336 B: one piece of a 64-row B tile
constexpr int H3_BTILE = 3 * H3_BPIECE;              // 43 008 B
constexpr int H3_LD = H3_BN + 4;                     // staging tile row stride (floats): conflict-free float4 rows
constexpr int H3_B_OFF = XTILE;
constexpr int H3_STAGE_OFF = H3_B_OFF + 2 * H3_BTILE;
constexpr int H3_SA_OFF = H3_STAGE_OFF + 128 * H3_LD * 4;
constexpr int H3_RED_OFF = H3_SA_OFF + 512;
constexpr int H3_BAR_OFF = H3_RED_OFF + 256;
constexpr int H3_SMEM = H3_BAR_OFF + 128;
constexpr uint32_t H3_SET_COLS = 3 * H3_BN;          // 192 TMEM columns per accumulator set (L0, L1, L2)

struct Head3Args {
  const unsigned char* a_img;   // [Cb, m_tiles, XTILE]       operand image of Bout (rows = functions)
  const float* a_sc;            // [Cb, m_tiles * 128]
  const unsigned char* b_img;   // [Cb, p_tiles128, XTILE]    operand image of Tout (rows = trunk points)
  const float* b_sc;            // [Cb, p_tiles128 * 128]
  int M, P, K;
  const float* bias;            // scalar output bias of chain c at bias[c * bias_bs]
  long long bias_bs;
  const float* Y;               // targets [M, ldy], shared by the chains (null: predict mode)
  long long ldy;
  float* G;                     // [Cb, M, ldg]: d loglik / d out (likelihood mode) or the outputs themselves (predict mode)
  long long g_bs, ldg;
  float ll_const, half_prec, prec;
  float* part_ll;               // [Cb, parts]
  float* part_g;
  int parts;                    // m_tiles * p_chunks
  int Cb, m_tiles, p_tiles128, p_tiles, p_chunks, tiles_per_chunk;
  int prefetch_ahead;           // B tiles pulled into L2 this many tiles ahead of the staged copy (0: none; measured: no effect)
  long long n_items;
};

// L2 residency hints: the targets are shared by every chain (40.8 MB: L2-resident if the 2.6 GB stream of G does not evict them;
// the first version re-read them from HBM for every chain: 2.58 GB of DRAM reads per 64-chain batch in ncu)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ld_hint(const float* p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_hint(float* p, const float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

template <bool PREDICT>
__global__ void __launch_bounds__(H3_THREADS, 1) head3_kernel(Head3Args a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* At = smem;
  float* stage = reinterpret_cast<float*>(smem + H3_STAGE_OFF);
  float* red = reinterpret_cast<float*>(smem + H3_RED_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + H3_BAR_OFF);
  uint64_t *a_full = bars, *a_empty = bars + 1, *b_full = bars + 2, *b_empty = bars + 4, *t_full = bars + 6, *t_empty = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tc::mbar_init(a_full, 1);
    tc::mbar_init(a_empty, 1);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(b_full + s, 1);
      tc::mbar_init(b_empty + s, 1);
      tc::mbar_init(t_full + s, 1);
      tc::mbar_init(t_empty + s, 16);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  // item -> (chain c fastest, so that concurrently running CTAs share the target tiles in L2; row tile mt; trunk-point chunk pc)
  auto decode = [&](long long item, int& c, int& mt, int& pc) {
    c = (int)(item % a.Cb);
    const long long rest = item / a.Cb;
    mt = (int)(rest % a.m_tiles);
    pc = (int)(rest / a.m_tiles);
  };

  if (warp == 0) {
    if (lane == 0) {   // ---------------- producer ----------------
      uint32_t it = 0, tcount = 0;
      for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        int c, mt, pc;
        decode(item, c, mt, pc);
        tc::mbar_wait(a_empty, (it & 1u) ^ 1u);
        expect_tx(a_full, (uint32_t)XTILE);
        bulk_load(At, a.a_img + ((long long)c * a.m_tiles + mt) * XTILE, (uint32_t)XTILE, a_full);
        const int pt_lo = pc * a.tiles_per_chunk, pt_hi = min(pt_lo + a.tiles_per_chunk, a.p_tiles);
        // Two shared-memory stages cannot hide the latency of a 43 KB tile that comes from HBM (the load of tile t+2 starts when the
        // MMAs of tile t finish and must land within the 0.7 us the MMAs of tile t+1 take): the tiles are pulled into L2 a few tiles
        // ahead, so the staged copy is an L2 hit.
        const int kAhead = a.prefetch_ahead;
        auto prefetch = [&](int pt) {
          const unsigned char* src = a.b_img + ((long long)c * a.p_tiles128 + (pt >> 1)) * XTILE + (pt & 1) * H3_BPIECE;
#pragma unroll
          for (int i = 0; i < 3; ++i)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (long long)i * XPIECE), "r"((uint32_t)H3_BPIECE) : "memory");
        };
        for (int pt = pt_lo; pt < pt_hi && pt < pt_lo + kAhead; ++pt) prefetch(pt);
        for (int pt = pt_lo; pt < pt_hi; ++pt, ++tcount) {
          const uint32_t st = tcount & 1u;
          if (pt + kAhead < pt_hi) prefetch(pt + kAhead);
          tc::mbar_wait(b_empty + st, ((tcount >> 1) & 1u) ^ 1u);
          expect_tx(b_full + st, (uint32_t)H3_BTILE);
          unsigned char* dst = smem + H3_B_OFF + st * H3_BTILE;
          const unsigned char* src = a.b_img + ((long long)c * a.p_tiles128 + (pt >> 1)) * XTILE + (pt & 1) * H3_BPIECE;
#pragma unroll
          for (int i = 0; i < 3; ++i) bulk_load(dst + i * H3_BPIECE, src + (long long)i * XPIECE, (uint32_t)H3_BPIECE, b_full + st);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---------------- MMA issuer ----------------
      uint32_t it = 0, tcount = 0;
      const int ksteps = (a.K + 15) / 16;
      for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        int c, mt, pc;
        decode(item, c, mt, pc);
        tc::mbar_wait(a_full, it & 1u);
        const int pt_lo = pc * a.tiles_per_chunk, pt_hi = min(pt_lo + a.tiles_per_chunk, a.p_tiles);
        for (int pt = pt_lo; pt < pt_hi; ++pt, ++tcount) {
          const uint32_t st = tcount & 1u, use = (tcount >> 1) & 1u;
          tc::mbar_wait(b_full + st, use);
          tc::mbar_wait(t_empty + st, use ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sb = tc::smem_u32(smem + H3_B_OFF + st * H3_BTILE);
          for (int ks = 0; ks < ksteps; ++ks)
            mma_step6(tmem_d + st * H3_SET_COLS, (uint32_t)H3_BN, tc::smem_u32(At), XPIECE, sb, H3_BPIECE, ks, idesc_bf16(H3_BN), ks == 0);
          tc::mma_commit(b_empty + st);   // the shared-memory stage is free once these MMAs have read it
          tc::mma_commit(t_full + st);    // ... and the accumulator set is complete
        }
        tc::mma_commit(a_empty);
      }
    }
  } else {             // ---------------- epilogue warps ----------------
    // 16 warps: four per TMEM lane quarter.  Phase 1 (thread = accumulator row, 16 columns): TMEM -> s_a (L0 + (L1 + L2)) -> staging
    // tile.  Phase 2 (thread = one float4 column group of 4 rows): coalesced rows of the targets in, rows of G out.  The ncu
    // profile of the first version (8 warps, 880 instructions per thread and tile, IPC 1.4, tensor pipe 26 %) showed the epilogue,
    // not the tensor core, setting the tile time: per-element masks, 64-bit address arithmetic and the per-element log-likelihood
    // arithmetic are gone from the inner loops (sums of res and res^2 are accumulated; masks only on edge tiles).
    const int e = warp - 2, q = warp & 3, colq = e >> 2;
    const int et = e * 32 + lane;                    // 0..511
    const int row1 = q * 32 + lane;
    // (measured and rejected: warp-private staging with no block barrier -- 8 rows x 64 B per store instruction instead of
    //  2 rows x 256 B: 1.52 ms against 1.33 ms; pulling the B tiles into L2 ahead of the staged copy: no change)
    const int c4 = et & 15, rr = et >> 4;            // phase 2: float4 column group c4 of rows rr + 32 i
    uint32_t tcount = 0;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    for (long long item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      int c, mt, pc;
      decode(item, c, mt, pc);
      const int m0 = mt * 128;
      // row scale of this thread's accumulator row, straight from global memory (the shared-memory copy of the scales would need
      // its own release protocol: an item of one or two tiles can be overtaken by the producer)
      const float sa_row = __ldg(a.a_sc + ((long long)c * a.m_tiles + mt) * 128 + row1);
      const float bias0 = __ldg(a.bias + (long long)c * a.bias_bs);
      float s1 = 0.0f, s2 = 0.0f;                    // sums of res and res^2 over this thread's elements
      int cnt = 0;
      const int pt_lo = pc * a.tiles_per_chunk, pt_hi = min(pt_lo + a.tiles_per_chunk, a.p_tiles);
      const bool rows_full = m0 + 128 <= a.M;
      const float* ybase = PREDICT ? nullptr : a.Y + (long long)(m0 + rr) * a.ldy + 4 * c4;
      float* gbase = a.G + (long long)c * a.g_bs + (long long)(m0 + rr) * a.ldg + 4 * c4;
      const float* sbbase = a.b_sc + (long long)c * a.p_tiles128 * 128 + 4 * c4;
      for (int pt = pt_lo; pt < pt_hi; ++pt, ++tcount) {
        const uint32_t st = tcount & 1u, use = (tcount >> 1) & 1u;
        const int p0 = pt * H3_BN + 4 * c4;          // this thread's first column
        const bool full = rows_full && (pt + 1) * H3_BN <= a.P;
        // targets and column scales of this thread's float4s, in flight while the tensor core works
        float4 yv[4];
        float4 sb4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p0 < a.P) sb4 = __ldg(reinterpret_cast<const float4*>(sbbase + pt * H3_BN));
        if (!PREDICT) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            yv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (full || (m0 + rr + 32 * i < a.M && p0 < a.P)) yv[i] = ld_hint(ybase + (long long)(32 * i) * a.ldy + pt * H3_BN, pol_keep);
          }
        }
        tc::mbar_wait(t_full + st, use);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {  // phase 1
          float* srow = stage + row1 * H3_LD + colq * 16;
          const uint32_t tbase = tmem_d + st * H3_SET_COLS + ((uint32_t)(q * 32) << 16) + (uint32_t)(colq * 16);
          uint32_t l0r[2][8], l1r[2][8], l2r[2][8];
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            tmem_ld8(tbase + (uint32_t)(g2 * 8), l0r[g2]);
            tmem_ld8(tbase + (uint32_t)(g2 * 8 + H3_BN), l1r[g2]);
            tmem_ld8(tbase + (uint32_t)(g2 * 8 + 2 * H3_BN), l2r[g2]);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const unsigned long long sa2 = tc::pk2(sa_row, sa_row);
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const unsigned long long low = add2(tc::pk2(__uint_as_float(l1r[g2][j]), __uint_as_float(l1r[g2][j + 1])),
                                                  tc::pk2(__uint_as_float(l2r[g2][j]), __uint_as_float(l2r[g2][j + 1])));
              const unsigned long long sum = add2(tc::pk2(__uint_as_float(l0r[g2][j]), __uint_as_float(l0r[g2][j + 1])), low);
              const unsigned long long sc = mul2(sum, sa2);
              asm("mov.b64 {%0, %1}, %2;" : "=f"(v[j]), "=f"(v[j + 1]) : "l"(sc));
            }
            *reinterpret_cast<float4*>(srow + g2 * 8) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(srow + g2 * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + st);    // the accumulator set may be overwritten
        named_bar_sync(1, 512);
        {  // phase 2
          const float* srow = stage + rr * H3_LD + 4 * c4;
          float* grow = gbase + pt * H3_BN;
          if (full) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 t = *reinterpret_cast<const float4*>(srow + 32 * i * H3_LD);
              const float o0 = fmaf(t.x, sb4.x, bias0), o1 = fmaf(t.y, sb4.y, bias0), o2 = fmaf(t.z, sb4.z, bias0), o3 = fmaf(t.w, sb4.w, bias0);
              if (PREDICT) {
                float* orow = grow + (long long)(32 * i) * a.ldg;
                orow[0] = o0; orow[1] = o1; orow[2] = o2; orow[3] = o3;
              } else {
                const float r0 = o0 - yv[i].x, r1 = o1 - yv[i].y, r2 = o2 - yv[i].z, r3 = o3 - yv[i].w;
                s1 += (r0 + r1) + (r2 + r3);
                s2 = fmaf(r0, r0, s2); s2 = fmaf(r1, r1, s2); s2 = fmaf(r2, r2, s2); s2 = fmaf(r3, r3, s2);
                st_hint(grow + (long long)(32 * i) * a.ldg, make_float4(-a.prec * r0, -a.prec * r1, -a.prec * r2, -a.prec * r3), pol_stream);
              }
            }
            cnt += 16;
          } else {
            const int nv = a.P - p0;                   // valid columns of this thread's float4
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int m = m0 + rr + 32 * i;
              if (m < a.M && nv > 0) {
                const float4 t = *reinterpret_cast<const float4*>(srow + 32 * i * H3_LD);
                const float o[4] = {fmaf(t.x, sb4.x, bias0), fmaf(t.y, sb4.y, bias0), fmaf(t.z, sb4.z, bias0), fmaf(t.w, sb4.w, bias0)};
                float* orow = grow + (long long)(32 * i) * a.ldg;
                if (PREDICT) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (j < nv) orow[j] = o[j];
                } else {
                  const float y[4] = {yv[i].x, yv[i].y, yv[i].z, yv[i].w};
                  float gq[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    gq[j] = 0.0f;
                    if (j < nv) {
                      const float res = o[j] - y[j];
                      s1 += res;
                      s2 = fmaf(res, res, s2);
                      gq[j] = -a.prec * res;
                      ++cnt;
                    }
                  }
                  *reinterpret_cast<float4*>(orow) = make_float4(gq[0], gq[1], gq[2], gq[3]);
                }
              }
            }
          }
        }
        named_bar_sync(1, 512);                      // the staging tile is free for the next tile
      }
      if (!PREDICT) {   // fixed-order sums of this item's partials: loglik = n ll_const - half_prec sum res^2, sum G = -prec sum res
        float ll_acc = fmaf(-a.half_prec, s2, (float)cnt * a.ll_const), g_acc = -a.prec * s1;
        ll_acc = warp_sum(ll_acc);
        g_acc = warp_sum(g_acc);
        if (lane == 0) { red[e] = ll_acc; red[16 + e] = g_acc; }
        named_bar_sync(1, 512);
        if (et == 0) {
          float t0 = 0.0f, t1 = 0.0f;
          for (int w = 0; w < 16; ++w) { t0 += red[w]; t1 += red[16 + w]; }
          a.part_ll[(long long)c * a.parts + mt * a.p_chunks + pc] = t0;
          a.part_g[(long long)c * a.parts + mt * a.p_chunks + pc] = t1;
        }
        named_bar_sync(1, 512);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

}  // namespace xg
}  // namespace vihmc
