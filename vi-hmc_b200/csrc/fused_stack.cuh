// All layers of one fully-connected stack for a 128-row tile in ONE kernel (forward pass of the DeepONet trunk / branch).
//
// Restates Operator_network/VI_HMC/my_make_func.py:53-61 (branch) and :69-77 (trunk): x = act(linear(x, W_l, b_l)) for every
// layer but the last, which has no activation.  The per-layer GEMM path (tc_gemm.cuh) moves every activation tile
// HBM -> registers -> split -> shared memory again for the next layer and spends most of its issue slots doing so; here the
// activations of a row tile never leave the SM between layers:
//   * the tile's activations live in shared memory as the 3xTF32 hi / lo operand pair (K-major, no swizzle, whole K = 104);
//   * layer l's weights arrive PRE-SPLIT in the same layout (weight_image_kernel builds the hi / lo images once per gradient
//     evaluation) with ONE cp.async.bulk + mbarrier::complete_tx per layer, overlapped with the previous layer's epilogue;
//   * one thread issues the 13 x 3 tcgen05.mma (M = 128, N = 112, K = 8) of a layer into TMEM (main + correction accumulators);
//   * the epilogue reads TMEM with the lane = row mapping, adds the bias, applies the activation, splits the result and
//     stores hi / lo straight into the operand tiles of the next layer (a quarter warp = 8 rows = 128 contiguous bytes:
//     conflict-free); the fp32 activations the backward pass needs are written to HBM by a coalesced pass over the
//     operand tiles (h = hi + lo, ~1 ulp from the unsplit value) that runs while the tensor core works on the next layer.
// Shared memory: 2 x 52 KB (activations) + 2 x 45.5 KB (weights) = 195 KB: one CTA per SM, 512 threads (the epilogue is a
// chain of TMEM load -> tanh -> split -> store latencies; 16 warps with a compile-time activation hide what 8 warps with a
// runtime one could not: 3.2 ms -> 1.5 ms for the 8 trunk layers of 64 chains, against 2.25 ms for the per-layer GEMMs).
// Eligibility (host side): every width a multiple of 4 and <= 104; an input of at most 8 features goes through an FP32 first
// layer (trunk), a wider one (<= 104, the branch's 101 sensors) is staged as the first tensor-core operand.
#pragma once
#include "tc_gemm.cuh"

namespace vihmc {
namespace fused {

using tc::BM;
constexpr int KPAD = 104, NPAD = 112;            // reduction / output widths the operand tiles are laid out for
constexpr int KCH = KPAD / 4;                     // 16-byte chunks per row
constexpr int RG_BYTES = KCH * 128;               // one 8-row group (the descriptors' SBO); LBO = 128
constexpr int A_TILE = (BM / 8) * RG_BYTES;       // 53,248 B
constexpr int B_TILE = (NPAD / 8) * RG_BYTES;     // 46,592 B
constexpr int F_SMEM = 2 * A_TILE + 2 * B_TILE + 128;
constexpr int F_THREADS = 512;                     // 16 warps: four per TMEM lane quarter, 32 accumulator columns each
#ifndef VIHMC_FUSED_DIRECT_STORE
#define VIHMC_FUSED_DIRECT_STORE 0   // 1: the epilogue writes fp32 activations from registers (row per thread: measured 1.73 ms
                                     // vs 1.49 ms); 0: coalesced pass over the operand tiles
#endif
constexpr int MAX_IN0 = 8;                        // widest first-layer input handled by the FP32 first layer
constexpr uint32_t kIdescF = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NPAD >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// ---- weight images ----
// Operand element (n, k) of image l is W[n*ld_n + k*ld_k] at w_off inside the chain's padded weight vector
// (forward: W[n][k]: ld_n = row stride, ld_k = 1).  Image = hi tile then lo tile, B_TILE bytes each, zero padded.
struct ImgLayer { long long w_off; int n_rows, n_k, ld_n, ld_k; };
struct ImgTable { int n; ImgLayer L[VIHMC_MAX_LAYERS]; };

__global__ void __launch_bounds__(256) weight_image_kernel(const float* __restrict__ Wf, long long Dp, ImgTable t, float* __restrict__ img) {
  const int l = blockIdx.x;
  const long long c = blockIdx.y;
  const ImgLayer L = t.L[l];
  const float* __restrict__ W = Wf + c * Dp + L.w_off;
  float* hi = img + (c * t.n + l) * (2 * B_TILE / 4);
  float* lo = hi + B_TILE / 4;
  for (int e = threadIdx.x; e < B_TILE / 4; e += blockDim.x) {
    const int j = e & 3, r8 = (e >> 2) & 7, ch = (e >> 5) % KCH, rg = (e >> 5) / KCH;
    const int n = rg * 8 + r8, k = ch * 4 + j;
    float v = 0.0f;
    if (n < L.n_rows && k < L.n_k) v = __ldg(W + (long long)n * L.ld_n + (long long)k * L.ld_k);
    const float h = tc::rna_tf32(v);
    hi[e] = h;
    lo[e] = tc::rna_tf32(v - h);
  }
}

struct FusedFwdArgs {
  const float* input;                     // [R, in_dim], shared by every chain
  int in_dim;
  const float* Wf;                        // padded weights [Cb, Dp]: first-layer weights and all biases are read from here
  long long Dp;
  long long w0_off;
  int ldw0;
  long long b_off[VIHMC_MAX_LAYERS];
  int dims[VIHMC_MAX_LAYERS];
  int n_layers;
  const float* img;                       // images of layers l0 .. n_layers-1 (l0 = 0 for a wide input, else 1): [Cb, n, 2, B_TILE/4]
  float* acts[VIHMC_MAX_LAYERS];          // acts[l]: [Cb, R, dims[l]]
  long long R;
  int act;                                // activation between layers (none after the last)
};

template <int ACT>
__device__ __forceinline__ float activate(float v) {
  if (ACT == VIHMC_ACT_TANH) return tanh_sel(v);
  if (ACT == VIHMC_ACT_RELU) return v > 0.0f ? v : 0.0f;
  return v;
}

template <int ACT>
__global__ void __launch_bounds__(F_THREADS, 1) fused_forward_kernel(FusedFwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A_hi = smem;
  unsigned char* A_lo = smem + A_TILE;
  unsigned char* B_hi = smem + 2 * A_TILE;          // the lo image follows contiguously, as in global memory
  uint64_t* bar_b = reinterpret_cast<uint64_t*>(smem + 2 * A_TILE + 2 * B_TILE);
  uint64_t* bar_mma = bar_b + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long c = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * BM;
  const float* __restrict__ Wc = a.Wf + c * a.Dp;
  // first layer: FP32 pipes when the input is narrow (trunk: 5 features), a tensor-core layer like the others when it is
  // wide (branch: 101 sensors) -- then the operand tiles start as the split input tile and images exist for every layer
  const int l0 = a.in_dim > MAX_IN0 ? 0 : 1;
  const int n_img = a.n_layers - l0;

  if (tid == 0) {
    tc::mbar_init(bar_b, 1);
    tc::mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the K padding of the operand tiles (chunks >= the first layer's width) must be zero: later layers only ever write
  // chunks below their own width and clear what a wider predecessor left
  for (int i = tid; i < BM * KCH; i += F_THREADS) {
    const int row = i % BM, ch = i / BM;
    if (l0 == 1 && 4 * ch >= a.dims[0]) {
      const int off = (row >> 3) * RG_BYTES + ch * 128 + (row & 7) * 16;
      *reinterpret_cast<float4*>(A_hi + off) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(A_lo + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  auto load_weights = [&](int l) {   // thread 0: one bulk copy of layer l's hi + lo images, completion on bar_b
    const float* src = a.img + (c * n_img + (l - l0)) * (2 * B_TILE / 4);
    const uint32_t bytes = 2u * B_TILE;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar_b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(B_hi)),
                 "l"(src), "r"(bytes), "r"(tc::smem_u32(bar_b))
                 : "memory");
  };
  if (tid == 0 && l0 < a.n_layers) load_weights(l0);

  if (l0 == 0) {
    // ---- wide input: the input tile itself becomes the first operand (rows of in_dim floats, any alignment) ----
    const int r8 = lane & 7, cq = lane >> 3, rg = warp;
    const long long grow = r0 + rg * 8 + r8;
    const float* __restrict__ src = a.input + grow * a.in_dim;
    for (int ch = cq; ch < KCH; ch += 4) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (grow < a.R && 4 * ch + j < a.in_dim) ? __ldg(src + 4 * ch + j) : 0.0f;
      float4 hi, lo;
      tc::split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
      const int off = rg * RG_BYTES + ch * 128 + r8 * 16;
      *reinterpret_cast<float4*>(A_hi + off) = hi;
      *reinterpret_cast<float4*>(A_lo + off) = lo;
    }
  } else {
    // ---- layer 0 on the FP32 pipes: warp -> row group, lane -> (row r8 = lane%8, chunks lane/8 + 4j) ----
    const int r8 = lane & 7, cq = lane >> 3, rg = warp;
    const int n0w = a.dims[0];
    const long long grow = r0 + rg * 8 + r8;
    float f[MAX_IN0];
#pragma unroll
    for (int k = 0; k < MAX_IN0; ++k) f[k] = (k < a.in_dim && grow < a.R) ? __ldg(a.input + grow * a.in_dim + k) : 0.0f;
    for (int ch = cq; 4 * ch < n0w; ch += 4) {
      float v[4];
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(Wc + a.b_off[0] + 4 * ch));
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = 4 * ch + j;
        const float4* wrow = reinterpret_cast<const float4*>(Wc + a.w0_off + (long long)n * a.ldw0);   // rows padded to 4 floats
        float w[MAX_IN0];
#pragma unroll
        for (int k4 = 0; k4 < MAX_IN0 / 4; ++k4) {
          const float4 t = 4 * k4 < a.ldw0 ? __ldg(wrow + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
          w[4 * k4] = t.x; w[4 * k4 + 1] = t.y; w[4 * k4 + 2] = t.z; w[4 * k4 + 3] = t.w;
        }
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < MAX_IN0; ++k)
          if (k < a.in_dim) acc = fmaf(w[k], f[k], acc);
        acc += bb[j];
        v[j] = a.n_layers > 1 ? activate<ACT>(acc) : acc;
      }
      if (VIHMC_FUSED_DIRECT_STORE && grow < a.R)
        *reinterpret_cast<float4*>(a.acts[0] + (c * a.R + grow) * n0w + 4 * ch) = make_float4(v[0], v[1], v[2], v[3]);
      float4 hi, lo;
      tc::split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
      const int off = rg * RG_BYTES + ch * 128 + r8 * 16;
      *reinterpret_cast<float4*>(A_hi + off) = hi;
      *reinterpret_cast<float4*>(A_lo + off) = lo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  // coalesced write of the activations held in the operand tiles: acts[l][c, r0 + row, :] = hi + lo
  auto store_acts = [&](int l) {
    const int r8 = lane & 7, cq = lane >> 3, rg = warp;
    const int nw = a.dims[l];
    const long long grow = r0 + rg * 8 + r8;
    if (grow >= a.R) return;
    float* __restrict__ out = a.acts[l] + (c * a.R + grow) * nw + 4 * cq;
    const unsigned char* ph = A_hi + rg * RG_BYTES + cq * 128 + r8 * 16;
    for (int ch = cq; 4 * ch < nw; ch += 4) {   // (unrolling this loop to batch the loads was measured: no gain)
      const float4 hi = *reinterpret_cast<const float4*>(ph);
      const float4 lo = *reinterpret_cast<const float4*>(ph + A_TILE);
      *reinterpret_cast<float4*>(out) = make_float4(hi.x + lo.x, hi.y + lo.y, hi.z + lo.z, hi.w + lo.w);
      ph += 4 * 128;
      out += 16;
    }
  };

  for (int l = l0; l < a.n_layers; ++l) {
    const int K = l == 0 ? a.in_dim : a.dims[l - 1], N = a.dims[l];
    const uint32_t ph = (uint32_t)(l - l0) & 1u;
    if (tid == 0) {
      tc::mbar_wait(bar_b, ph);   // this layer's weight images have landed
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = tc::smem_u32(A_hi), sb = tc::smem_u32(B_hi);
      const int ksteps = (K + 7) / 8;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t ko = (uint32_t)ks * 256u;   // one K = 8 step = two 16-byte chunks, 128 B apart
        const uint64_t a_hi = tc::make_desc(sa + ko, 128, RG_BYTES), a_lo = tc::make_desc(sa + A_TILE + ko, 128, RG_BYTES);
        const uint64_t b_hi = tc::make_desc(sb + ko, 128, RG_BYTES), b_lo = tc::make_desc(sb + B_TILE + ko, 128, RG_BYTES);
        const uint32_t acc = ks > 0 ? 1u : 0u;
        tc::mma_tf32(tmem_d, a_hi, b_hi, kIdescF, acc);            // main product
        tc::mma_tf32(tmem_d + 128, a_lo, b_hi, kIdescF, acc);      // corrections in their own accumulator (see tc_gemm.cuh)
        tc::mma_tf32(tmem_d + 128, a_hi, b_lo, kIdescF, 1u);
      }
      tc::mma_commit(bar_mma);
    }
    if (!VIHMC_FUSED_DIRECT_STORE && l >= 1) store_acts(l - 1);   // reads the operand tiles while the tensor core reads them too
    // this thread's 32 bias values, fetched while the tensor core works (they were the first consumer after the TMEM loads
    // of every 8-column group: four exposed L2 latencies per layer)
    float4 bpre[8];
    {
      const float* __restrict__ bias = Wc + a.b_off[l];
      const int nb = (warp >> 2) * 32;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        bpre[i] = nb + 4 * i < N ? __ldg(reinterpret_cast<const float4*>(bias + nb + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc::mbar_wait(bar_mma, ph);   // accumulators complete; operand and weight tiles are free
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();              // every thread is done reading the operand tiles (store_acts)
    if (tid == 0 && l + 1 < a.n_layers) load_weights(l + 1);
    {  // epilogue: TMEM (lane = row) -> bias, activation -> split -> operand tiles of the next layer
      const int q = warp & 3, cq4 = warp >> 2;          // TMEM lane quarter; 32-column quarter of the accumulator
      const int row = q * 32 + lane;
      unsigned char* prow = A_hi + (row >> 3) * RG_BYTES + (row & 7) * 16;
      const bool last = l == a.n_layers - 1;
      const long long grow = r0 + row;
      float* __restrict__ orow = a.acts[l] + (c * a.R + grow) * N;
      // two 16-column halves: the four TMEM loads (main + correction accumulator of two 8-column groups) of a half are in
      // flight together behind ONE tcgen05.wait::ld instead of one wait per group
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int nh = cq4 * 32 + hh * 16;
        if (nh >= N) break;
        uint32_t r[2][8], rc[2][8];
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(nh + 8 * g2);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[g2][0]), "=r"(r[g2][1]), "=r"(r[g2][2]), "=r"(r[g2][3]), "=r"(r[g2][4]), "=r"(r[g2][5]), "=r"(r[g2][6]),
                         "=r"(r[g2][7])
                       : "r"(taddr));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(rc[g2][0]), "=r"(rc[g2][1]), "=r"(rc[g2][2]), "=r"(rc[g2][3]), "=r"(rc[g2][4]), "=r"(rc[g2][5]),
                         "=r"(rc[g2][6]), "=r"(rc[g2][7])
                       : "r"(taddr + 128u));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const int n = nh + 8 * g2, cc = hh * 16 + 8 * g2;
          if (n >= N) break;
          const bool two = n + 4 < N;   // widths are multiples of 4: the second float4 of the group is all in or all out
          const float4 b0 = bpre[cc >> 2], b1 = bpre[(cc >> 2) + 1];
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = __uint_as_float(r[g2][j]) + __uint_as_float(rc[g2][j]) + bb[j];
            v[j] = last ? x : activate<ACT>(x);
          }
          if (VIHMC_FUSED_DIRECT_STORE && grow < a.R) {
            *reinterpret_cast<float4*>(orow + n) = make_float4(v[0], v[1], v[2], v[3]);
            if (two) *reinterpret_cast<float4*>(orow + n + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          float4 hi, lo;
          tc::split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
          *reinterpret_cast<float4*>(prow + (n >> 2) * 128) = hi;
          *reinterpret_cast<float4*>(prow + A_TILE + (n >> 2) * 128) = lo;
          if (two) {
            tc::split4(make_float4(v[4], v[5], v[6], v[7]), hi, lo);
            *reinterpret_cast<float4*>(prow + ((n >> 2) + 1) * 128) = hi;
            *reinterpret_cast<float4*>(prow + A_TILE + ((n >> 2) + 1) * 128) = lo;
          }
        }
      }
      // a narrower layer leaves stale columns [N, K) of the previous one in the tiles: clear them (never taken at equal widths)
      for (int ch = N / 4 + cq4; ch < (K + 3) / 4; ch += 4) {
        *reinterpret_cast<float4*>(prow + ch * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(prow + A_TILE + ch * 128) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (!VIHMC_FUSED_DIRECT_STORE) store_acts(a.n_layers - 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// Backward data pass of the same stack: dz_{l-1} = (dz_l W_l) * act'(a_{l-1}) for l = top .. 1, one kernel per row tile.
// (autograd of my_make_func.py:53-61 / :69-77).  Same machinery as the forward kernel with the transposed weight images
// (operand element (n = input unit, k = output unit) = W_l[k][n]); per layer:
//   MMA (dz_l in the operand tiles x W_l^T image) -> TMEM; the epilogue stores the raw product as fp32 into the hi tile;
//   a coalesced pass multiplies by act'(a_{l-1}) (activation tile prefetched from HBM into registers while the tensor core
//   works), writes dz_{l-1} to HBM for the weight-gradient GEMM and splits it into the operand tiles of the next layer.
// ---------------------------------------------------------------------------------------------------------------
struct FusedBwdArgs {
  int dims[VIHMC_MAX_LAYERS];
  int n_layers;
  const float* img;                        // transposed images of layers 1 .. n_layers-1: [Cb, n_layers-1, 2, B_TILE/4]
  const float* acts[VIHMC_MAX_LAYERS];     // acts[l]: [Cb, R, dims[l]] (forward activations)
  float* dz[VIHMC_MAX_LAYERS];             // dz[l]: [Cb, R, dims[l]]; dz[n_layers-1] is the input, the others are written
  long long R;
};

template <int ACT>
__global__ void __launch_bounds__(F_THREADS, 1) fused_backward_kernel(FusedBwdArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A_hi = smem;
  unsigned char* A_lo = smem + A_TILE;
  unsigned char* B_hi = smem + 2 * A_TILE;
  uint64_t* bar_b = reinterpret_cast<uint64_t*>(smem + 2 * A_TILE + 2 * B_TILE);
  uint64_t* bar_mma = bar_b + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_b + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long c = blockIdx.y;
  const long long r0 = (long long)blockIdx.x * BM;
  const int n_img = a.n_layers - 1, top = a.n_layers - 1;

  if (tid == 0) {
    tc::mbar_init(bar_b, 1);
    tc::mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  auto load_weights = [&](int l) {   // thread 0: image of layer l (images are stored in layer order 1 .. top)
    const float* src = a.img + (c * n_img + (l - 1)) * (2 * B_TILE / 4);
    const uint32_t bytes = 2u * B_TILE;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar_b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(B_hi)),
                 "l"(src), "r"(bytes), "r"(tc::smem_u32(bar_b))
                 : "memory");
  };
  if (tid == 0) load_weights(top);

  // coalesced mapping of the passes: warp -> row group, lane -> (row r8 = lane%8, chunks lane/8 + 4i)
  const int r8 = lane & 7, cq = lane >> 3, rg = warp;
  const long long grow = r0 + rg * 8 + r8;
  const bool rvalid = grow < a.R;
  const int poff = rg * RG_BYTES + cq * 128 + r8 * 16;
  constexpr int NCH = (KCH + 3) / 4;

  {  // incoming gradient tile -> operand tiles (zero beyond its width and beyond the matrix)
    const int nw = a.dims[top];
    const float* __restrict__ src = a.dz[top] + (c * a.R + grow) * nw + 4 * cq;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const int ch = cq + 4 * i;
      if (ch >= KCH) break;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rvalid && 4 * ch < nw) v = __ldg(reinterpret_cast<const float4*>(src + 16 * i));
      float4 hi, lo;
      tc::split4(v, hi, lo);
      *reinterpret_cast<float4*>(A_hi + poff + i * 4 * 128) = hi;
      *reinterpret_cast<float4*>(A_lo + poff + i * 4 * 128) = lo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  uint32_t ph = 0;
  for (int l = top; l >= 1; --l, ph ^= 1u) {
    const int K = a.dims[l], N = a.dims[l - 1];
    if (tid == 0) {
      tc::mbar_wait(bar_b, ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = tc::smem_u32(A_hi), sb = tc::smem_u32(B_hi);
      const int ksteps = (K + 7) / 8;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t ko = (uint32_t)ks * 256u;
        const uint64_t a_hi = tc::make_desc(sa + ko, 128, RG_BYTES), a_lo = tc::make_desc(sa + A_TILE + ko, 128, RG_BYTES);
        const uint64_t b_hi = tc::make_desc(sb + ko, 128, RG_BYTES), b_lo = tc::make_desc(sb + B_TILE + ko, 128, RG_BYTES);
        const uint32_t acc = ks > 0 ? 1u : 0u;
        tc::mma_tf32(tmem_d, a_hi, b_hi, kIdescF, acc);
        tc::mma_tf32(tmem_d + 128, a_lo, b_hi, kIdescF, acc);
        tc::mma_tf32(tmem_d + 128, a_hi, b_lo, kIdescF, 1u);
      }
      tc::mma_commit(bar_mma);
    }
    // activation tile of layer l-1 into registers while the tensor core works
    float4 h[NCH];
    {
      const float* __restrict__ hsrc = a.acts[l - 1] + (c * a.R + grow) * N + 4 * cq;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        h[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rvalid && 4 * (cq + 4 * i) < N) h[i] = __ldg(reinterpret_cast<const float4*>(hsrc + 16 * i));
      }
    }
    tc::mbar_wait(bar_mma, ph);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && l - 1 >= 1) load_weights(l - 1);   // the weight tiles are free: fetch the next image
    {  // epilogue: raw product (main + corrections) as fp32 into the hi tile, TMEM lane = row
      const int q = warp & 3, cq4 = warp >> 2;
      const int row = q * 32 + lane;
      unsigned char* prow = A_hi + (row >> 3) * RG_BYTES + (row & 7) * 16;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {   // two 16-column halves, four TMEM loads behind one wait (as in the forward kernel)
        const int nh = cq4 * 32 + hh * 16;
        if (nh >= N) break;
        uint32_t r[2][8], rc[2][8];
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(nh + 8 * g2);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[g2][0]), "=r"(r[g2][1]), "=r"(r[g2][2]), "=r"(r[g2][3]), "=r"(r[g2][4]), "=r"(r[g2][5]), "=r"(r[g2][6]),
                         "=r"(r[g2][7])
                       : "r"(taddr));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(rc[g2][0]), "=r"(rc[g2][1]), "=r"(rc[g2][2]), "=r"(rc[g2][3]), "=r"(rc[g2][4]), "=r"(rc[g2][5]),
                         "=r"(rc[g2][6]), "=r"(rc[g2][7])
                       : "r"(taddr + 128u));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int g2 = 0; g2 < 2; ++g2) {
          const int n = nh + 8 * g2;
          if (n >= N) break;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g2][j]) + __uint_as_float(rc[g2][j]);
          *reinterpret_cast<float4*>(prow + (n >> 2) * 128) = make_float4(v[0], v[1], v[2], v[3]);
          if (n + 4 < N) *reinterpret_cast<float4*>(prow + ((n >> 2) + 1) * 128) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    {  // pass: dz_{l-1} = raw * act'(a_{l-1}); HBM copy for the weight-gradient GEMM; hi / lo operands of the next layer
      float* __restrict__ out = a.dz[l - 1] + (c * a.R + grow) * N + 4 * cq;
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        const int ch = cq + 4 * i;
        if (ch >= KCH) break;
        unsigned char* pa = A_hi + poff + i * 4 * 128;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);   // chunks in [N, K) held the previous layer's operand: cleared
        if (4 * ch < N) {
          const float4 raw = *reinterpret_cast<const float4*>(pa);
          if (ACT == VIHMC_ACT_TANH) {
            v.x = raw.x * (1.0f - h[i].x * h[i].x); v.y = raw.y * (1.0f - h[i].y * h[i].y);
            v.z = raw.z * (1.0f - h[i].z * h[i].z); v.w = raw.w * (1.0f - h[i].w * h[i].w);
          } else {
            v.x = h[i].x > 0.0f ? raw.x : 0.0f; v.y = h[i].y > 0.0f ? raw.y : 0.0f;
            v.z = h[i].z > 0.0f ? raw.z : 0.0f; v.w = h[i].w > 0.0f ? raw.w : 0.0f;
          }
          if (rvalid) *reinterpret_cast<float4*>(out + 16 * i) = v;
        } else if (4 * ch >= K) {
          continue;   // beyond both widths: already zero
        }
        float4 hi, lo;
        tc::split4(v, hi, lo);
        *reinterpret_cast<float4*>(pa) = hi;
        *reinterpret_cast<float4*>(pa + A_TILE) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
}

}  // namespace fused
}  // namespace vihmc
