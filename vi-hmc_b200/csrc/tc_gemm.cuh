// Tensor-core batched GEMM for the dense path: tcgen05.mma kind::tf32 with FP32 accumulators in TMEM,
// 3xTF32 operand splitting for fp32-grade accuracy, and the dense path's fused epilogues.
//
//   C[b] = opA(A[b]) (MxK) * opB(B[b]) (KxN)          (same GemmArgs as the FP32-SIMT kernel in dense.cu)
//
// Why 3xTF32: the parity bar is rtol 1e-5 against the reference's fp32 closure; one TF32 product carries
// ~1e-3.  Every fp32 operand v is split into hi = rna_tf32(v) and lo = rna_tf32(v - hi) (both written to
// shared memory as exact tf32 values, low 13 mantissa bits zero, so the result does not depend on how the
// tensor core converts fp32 bits to tf32) and the product is hi*hi + lo*hi + hi*lo, accumulated in fp32.
//
// Structure of one CTA (256 threads, one 128 x 128 output tile, 3 x 32 KB smem stages, 256 TMEM columns => exactly
// 2 CTAs / SM, so no third CTA sits spinning in tcgen05.alloc so one CTA's
// epilogue overlaps another's main loop):
//   * all 8 warps stream the A / B k-tiles (16 floats deep) from global memory through registers, split
//     them, and store hi/lo tiles in the UMMA canonical K-major no-swizzle layout (8x16B core matrices);
//     arbitrary element strides are supported, so transposed operands (dW = dZ^T X, dX = dZ W) and
//     rows that are not 16-byte aligned (101-wide branch input) need no extra copies -- this is also
//     why the operands are not staged with TMA: the split needs the values in registers anyway.
//   * three smem stages; thread 0 issues 3 MMAs per 8-deep k-step and commits to the stage's mbarrier;
//     the next tile's global loads are in flight while the tensor core works.
//   * epilogue: 8 warps read the accumulators with tcgen05.ld (warp w -> TMEM lanes 32*(w%4).., column
//     half w/4) into a shared-memory tile, then apply bias+act / act' / Gaussian residual and write C
//     row-wise so every warp instruction touches contiguous memory.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace vihmc {

enum Epilogue {
  EPI_STORE = 0,      // C = acc
  EPI_BIAS_ACT = 1,   // C = act(acc + bias[n])           (act = identity when act < 0)
  EPI_DACT = 2,       // C = acc * act'(aux[m,n])         (aux = the layer's stored activation)
  EPI_HEAD = 3        // r = acc + bias0 - Y[m,n]; C = -prec r; partial sums of loglik and of C per CTA
};

struct GemmArgs {
  const float* A; long long a_bs, a_sm, a_sk;
  const float* B; long long b_bs, b_sk, b_sn;
  float* C; long long c_bs, ldc;
  int M, N, K;
  // epilogue operands
  const float* bias; long long bias_bs;   // [N] per batch (EPI_BIAS_ACT) or scalar per batch (EPI_HEAD)
  const float* aux; long long aux_bs, ld_aux;  // activation (EPI_DACT) or Y (EPI_HEAD; aux_bs = 0: shared)
  int act;
  float ll_const, half_prec, prec;
  float* part_ll; float* part_g;          // [batch, tiles] (EPI_HEAD)
  // split-K (EPI_STORE only): blockIdx.z = b * splits + s handles k in [s*kc, min(K,(s+1)*kc)) and writes its
  // partial product to split_buf[(s*batch + b), M, N]; splits <= 1 means a plain GEMM
  int splits, kc, batch;
  float* split_buf;
};

namespace tc {

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, STAGES = 3;
constexpr int CHUNKS = BK / 4;                 // 16-byte chunks along K per row
constexpr int LBO = 128;                       // bytes between the K-chunks of a core-matrix row group
constexpr int SBO = CHUNKS * 128;              // bytes between 8-row groups
constexpr int TILE_BYTES = (BM / 8) * SBO;     // 8 KB: one 128 x 16 fp32 operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;    // A_hi, A_lo, B_hi, B_lo
constexpr int TILE_LD = BN + 4;                 // epilogue staging tile row stride (floats): conflict-free float4 rows
constexpr int EPI_BYTES = BM * TILE_LD * 4;     // 67,584 B >= the two operand stages
constexpr int SMEM_BYTES = (EPI_BYTES > STAGES * STAGE_BYTES ? EPI_BYTES : STAGES * STAGE_BYTES) + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | (1ull << 46);
}

// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = BN (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc),
      "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a wedged barrier traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}

// 3xTF32 split with ROUND-TO-NEAREST: hi = rna_tf32(v), lo = rna_tf32(v - hi) (v - hi is exact in fp32).
// Truncating instead (v & 0xFFFFE000) shrinks every operand toward zero, i.e. a coherent ~1e-6 relative
// bias on every output; sums with heavy cancellation (d/d b0 = sum of G over N*P outputs) then miss the
// 1e-5 parity bar.  With rounding the per-product error is zero-mean.
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  hi.x = rna_tf32(v.x); lo.x = rna_tf32(v.x - hi.x);
  hi.y = rna_tf32(v.y); lo.y = rna_tf32(v.y - hi.y);
  hi.z = rna_tf32(v.z); lo.z = rna_tf32(v.z - hi.z);
  hi.w = rna_tf32(v.w); lo.w = rna_tf32(v.w - hi.w);
}

// One operand tile: rows [r0, r0+128) x k [k0, k0+16).  Element (r, k) lives at base + r*s_r + k*s_k.
// A warp instruction covers one 8-row group x 4 chunks: lane -> (row r8 = lane%8, chunk = lane/8), which
// makes the 16-byte shared-memory stores conflict-free and reads 64 contiguous bytes per row.
struct TileLoader {
  const float* base;
  long long s_r, s_k;
  int R, K, r0;
  bool vec;   // K-contiguous and 16-byte aligned: float4 loads
  __device__ __forceinline__ void fetch(int k0, int tid, float4 (&v)[2]) const {
    const int lane = tid & 31, warp = tid >> 5;
    const int r8 = lane & 7, ch = lane >> 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + (i * 8 + warp) * 8 + r8;
      const int k = k0 + ch * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < R) {
        const float* p = base + (long long)r * s_r + (long long)k * s_k;
        if (vec) {
          if (k < K) x = __ldg(reinterpret_cast<const float4*>(p));   // K % 4 == 0 on this path
        } else {
          if (k + 0 < K) x.x = __ldg(p);
          if (k + 1 < K) x.y = __ldg(p + s_k);
          if (k + 2 < K) x.z = __ldg(p + 2 * s_k);
          if (k + 3 < K) x.w = __ldg(p + 3 * s_k);
        }
      }
      v[i] = x;
    }
  }
};

__device__ __forceinline__ void stash(unsigned char* hi_tile, unsigned char* lo_tile, int tid, const float4 (&v)[2]) {
  const int lane = tid & 31, warp = tid >> 5;
  const int r8 = lane & 7, ch = lane >> 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int off = (i * 8 + warp) * SBO + ch * LBO + r8 * 16;
    float4 hi, lo;
    split4(v[i], hi, lo);
    *reinterpret_cast<float4*>(hi_tile + off) = hi;
    *reinterpret_cast<float4*>(lo_tile + off) = lo;
  }
}


// Phase-2 epilogue of one staged row: applies the fused epilogue to tile_row[0..BN) and writes C.  With `vec`
// (C rows 16-byte aligned, N % 4 == 0) every lane handles 4 consecutive columns with one LDS.128 / STG.128
// (one warp instruction = one 512-byte row); otherwise lane + 32*j columns (128 contiguous bytes each).
template <int EPI>
__device__ __forceinline__ void epilogue_row(const GemmArgs& g, const float* tile_row, float* crow, const float* arow,
                                             const float* biasb, int n0, int lane, bool vec, bool vec_aux, float bias0,
                                             float& ll_acc, float& g_acc) {
  auto apply = [&](float v, float bias_v, float aux_v) -> float {
    if (EPI == EPI_BIAS_ACT) {
      v += bias_v;
      if (g.act == VIHMC_ACT_TANH) v = tanh_sel(v);   // 8 instructions, abs. error ~1.2e-7 (see common.cuh)
      else if (g.act == VIHMC_ACT_RELU) v = v > 0.0f ? v : 0.0f;
    } else if (EPI == EPI_DACT) {
      v *= (g.act == VIHMC_ACT_TANH) ? (1.0f - aux_v * aux_v) : (aux_v > 0.0f ? 1.0f : 0.0f);
    } else if (EPI == EPI_HEAD) {
      const float res = v + bias0 - aux_v;
      ll_acc += g.ll_const - g.half_prec * res * res;
      v = -g.prec * res;
      g_acc += v;
    }
    return v;
  };
  if (vec) {
    const int c = lane * 4;
    if (n0 + c < g.N) {   // N % 4 == 0: the whole float4 is in range
      const float4 v = *reinterpret_cast<const float4*>(tile_row + c);
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), av = make_float4(0.f, 0.f, 0.f, 0.f);
      if (EPI == EPI_BIAS_ACT) {
        bv.x = __ldg(biasb + n0 + c); bv.y = __ldg(biasb + n0 + c + 1); bv.z = __ldg(biasb + n0 + c + 2); bv.w = __ldg(biasb + n0 + c + 3);
      }
      if (EPI == EPI_DACT || EPI == EPI_HEAD) {
        if (vec_aux) av = __ldg(reinterpret_cast<const float4*>(arow + c));
        else { av.x = __ldg(arow + c); av.y = __ldg(arow + c + 1); av.z = __ldg(arow + c + 2); av.w = __ldg(arow + c + 3); }
      }
      float4 o;
      o.x = apply(v.x, bv.x, av.x); o.y = apply(v.y, bv.y, av.y); o.z = apply(v.z, bv.z, av.z); o.w = apply(v.w, bv.w, av.w);
      *reinterpret_cast<float4*>(crow + c) = o;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      if (n0 + c >= g.N) continue;
      const float bias_v = (EPI == EPI_BIAS_ACT) ? __ldg(biasb + n0 + c) : 0.0f;
      const float aux_v = (EPI == EPI_DACT || EPI == EPI_HEAD) ? __ldg(arow + c) : 0.0f;
      crow[c] = apply(tile_row[c], bias_v, aux_v);
    }
  }
}

__device__ __forceinline__ bool rows_vec_ok(const float* base, long long bs, long long ld, int N) {
  return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && bs % 4 == 0 && ld % 4 == 0 && N % 4 == 0;
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 2) tc_gemm_kernel(GemmArgs g, int a_vec, int b_vec) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + SMEM_BYTES - 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SMEM_BYTES - 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool split = (EPI == EPI_STORE) && g.splits > 1;
  const int b = split ? (int)blockIdx.z / g.splits : (int)blockIdx.z;
  const int ksplit = split ? (int)blockIdx.z % g.splits : 0;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (split) {   // restrict this CTA to its K slice and redirect the output to the partial buffer
    const int k_lo = ksplit * g.kc;
    g.A += (long long)k_lo * g.a_sk;
    g.B += (long long)k_lo * g.b_sk;
    g.K = (g.K - k_lo) < g.kc ? (g.K - k_lo) : g.kc;
    g.C = g.split_buf + ((long long)ksplit * g.batch) * (long long)g.M * g.N;
    g.c_bs = (long long)g.M * g.N;
    g.ldc = g.N;
  }

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&mbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;

  TileLoader la{g.A + (long long)b * g.a_bs, g.a_sm, g.a_sk, g.M, g.K, m0, a_vec != 0};
  TileLoader lb{g.B + (long long)b * g.b_bs, g.b_sn, g.b_sk, g.N, g.K, n0, b_vec != 0};

  const int nk = (g.K + BK - 1) / BK;
  float4 ra[2], rb[2];
  la.fetch(0, tid, ra);
  lb.fetch(0, tid, rb);
  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt % STAGES;
    unsigned char* st = smem + s * STAGE_BYTES;
    if (kt >= STAGES) mbar_wait(&mbar[s], (uint32_t)(kt / STAGES - 1) & 1u);   // the MMAs that read this stage are done
    stash(st, st + TILE_BYTES, tid, ra);
    stash(st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, tid, rb);
    if (kt + 1 < nk) {   // next tile's global loads are in flight while the tensor core works on this one
      la.fetch((kt + 1) * BK, tid, ra);
      lb.fetch((kt + 1) * BK, tid, rb);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa = smem_u32(st);
      const int k_left = g.K - kt * BK;
      const int steps = k_left > 8 ? 2 : 1;
      for (int ks = 0; ks < steps; ++ks) {
        const uint32_t koff = (uint32_t)ks * 2u * LBO;   // one K=8 step = two 16-byte chunks
        const uint64_t a_hi = make_desc(sa + koff), a_lo = make_desc(sa + TILE_BYTES + koff);
        const uint64_t b_hi = make_desc(sa + 2 * TILE_BYTES + koff), b_lo = make_desc(sa + 3 * TILE_BYTES + koff);
        // The tensor core's fp32 accumulate rounds toward zero: every accumulate step shrinks the running sum
        // by up to one ulp, a COHERENT bias that grows with the number of chained MMAs.  The two correction
        // products therefore get their own accumulator (columns BN..2BN): the main chain sees one accumulate
        // per k-step instead of three, and the correction chain's truncation is 2^-11 smaller.
        const uint32_t acc = (kt > 0 || ks > 0) ? 1u : 0u;
        mma_tf32(tmem_d, a_hi, b_hi, acc);
        mma_tf32(tmem_d + BN, a_lo, b_hi, acc);
        mma_tf32(tmem_d + BN, a_hi, b_lo, 1u);
      }
      mma_commit(&mbar[s]);
    }
  }
  mbar_wait(&mbar[(nk - 1) % STAGES], (uint32_t)((nk - 1) / STAGES) & 1u);   // commits complete in order: everything is done
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---------------- epilogue ----------------
  // phase 1: TMEM -> registers -> shared tile [128][TILE_LD] (thread = one accumulator row, 8 columns per ld).
  // phase 2: coalesced pass: warp w owns rows w, w+8, ...; a warp instruction touches 512 contiguous bytes of
  // C / aux (each lane 4 columns), instead of 32 different rows as a row-per-thread epilogue would.
  float* tile = reinterpret_cast<float*>(smem);   // operand stages are dead: every MMA has completed
  {
    const int q = warp & 3, half = warp >> 2;
    float* trow = tile + (q * 32 + lane) * TILE_LD + half * (BN / 2);
#pragma unroll 2
    for (int cc = 0; cc < BN / 2; cc += 8) {
      uint32_t r[8], rc[8];
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2) + cc);
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr));
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(rc[0]), "=r"(rc[1]), "=r"(rc[2]), "=r"(rc[3]), "=r"(rc[4]), "=r"(rc[5]), "=r"(rc[6]), "=r"(rc[7])
                   : "r"(taddr + (uint32_t)BN));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(rc[j]);   // main + corrections (RN)
      *reinterpret_cast<float4*>(trow + cc) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(trow + cc + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
  __syncthreads();
  float* __restrict__ Cb = g.C + (long long)b * g.c_bs;
  const float bias0 = (EPI == EPI_HEAD) ? __ldg(g.bias + (long long)b * g.bias_bs) : 0.0f;
  const float* auxb = (EPI == EPI_DACT || EPI == EPI_HEAD) ? g.aux + (long long)b * g.aux_bs : nullptr;
  const float* biasb = (EPI == EPI_BIAS_ACT) ? g.bias + (long long)b * g.bias_bs : nullptr;
  const bool vec = rows_vec_ok(g.C, g.c_bs, g.ldc, g.N);
  const bool vec_aux = auxb != nullptr && rows_vec_ok(g.aux, g.aux_bs, g.ld_aux, g.N);
  float ll_acc = 0.0f, g_acc = 0.0f;
#pragma unroll 2
  for (int r = warp; r < BM; r += THREADS / 32) {
    const int m = m0 + r;
    if (m >= g.M) break;
    epilogue_row<EPI>(g, tile + r * TILE_LD, Cb + (long long)m * g.ldc + n0, auxb ? auxb + (long long)m * g.ld_aux + n0 : nullptr,
                      biasb, n0, lane, vec, vec_aux, bias0, ll_acc, g_acc);
  }
  if (EPI == EPI_HEAD) {
    float* red = reinterpret_cast<float*>(smem);
    ll_acc = warp_sum(ll_acc);
    g_acc = warp_sum(g_acc);
    __syncthreads();   // every warp is done reading the staged tile
    if (lane == 0) { red[warp] = ll_acc; red[8 + warp] = g_acc; }
    __syncthreads();
    if (tid == 0) {
      float s0 = 0.0f, s1 = 0.0f;
      for (int w = 0; w < THREADS / 32; ++w) { s0 += red[w]; s1 += red[8 + w]; }
      const long long tiles = (long long)gridDim.x * gridDim.y;
      const long long t = (long long)blockIdx.y * gridDim.x + blockIdx.x;
      g.part_ll[(long long)b * tiles + t] = s0;
      g.part_g[(long long)b * tiles + t] = s1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)(2 * BN)) : "memory");
}

}  // namespace tc

// =============================================================================================
// Persistent, warp-specialised variant (opt-in, VIHMC_TC_PERSISTENT=1): one CTA per SM loops over output tiles.
//   warps 0-7   producers: global -> registers -> 3xTF32 split -> shared-memory stage (4-stage ring, full/empty
//               mbarriers), running ahead of the tensor core by up to four k-tiles, across tile boundaries;
//   warp  8     MMA issuer: waits full[s], issues the three tcgen05.mma of each k-step into accumulator set
//               a = tile & 1 (TMEM columns a*256 .. a*256+255: main + correction), tcgen05.commit -> empty[s];
//               after the last k-tile tcgen05.commit -> tmem_full[a];
//   warps 9-16  epilogue (two per TMEM lane quarter, 64 columns each): wait tmem_full[a], tcgen05.ld the 128x128 accumulators into a shared staging tile,
//               release the accumulator set (tmem_empty[a]) so the MMAs of tile i+2 can start, then the
//               row-wise coalesced epilogue of tile i runs while the tensor core works on tile i+1.
// Compared with the one-tile-per-CTA kernel this removes the per-tile TMEM allocation / barrier setup and
// overlaps epilogue and main loop inside one SM (K = 100 layers have only 7 k-tiles per tile).
// =============================================================================================
namespace tcws {

using namespace tc;
constexpr int WS_STAGES = 4;
constexpr int PRODUCER_WARPS = 8, EPI_WARPS = 8;
constexpr int WS_THREADS = (PRODUCER_WARPS + 1 + EPI_WARPS) * 32;   // 544
constexpr int WS_EPI_OFF = WS_STAGES * STAGE_BYTES;                  // 128 KB of operand stages
constexpr int WS_BAR_OFF = WS_EPI_OFF + EPI_BYTES;                   // + 67,584 B staging tile
constexpr int WS_SMEM_BYTES = WS_BAR_OFF + 256;
constexpr int EPI_BAR_ID = 1;                                        // named barrier of the 128 epilogue threads

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync %0, %1;" ::"n"(EPI_BAR_ID), "n"(EPI_WARPS * 32) : "memory"); }

struct TileCoord {
  int b, m0, n0, K;
  const float *A, *B;
  float* C;
  long long c_bs, ldc;
  long long tile_linear;   // index into the per-batch partial arrays (EPI_HEAD)
};

__device__ __forceinline__ TileCoord decode_tile(const GemmArgs& g, long long t, int tiles_n, int tiles_m) {
  TileCoord tc;
  const int bx = (int)(t % tiles_n), by = (int)((t / tiles_n) % tiles_m), bz = (int)(t / ((long long)tiles_n * tiles_m));
  const bool split = g.splits > 1;
  tc.b = split ? bz / g.splits : bz;
  const int ks = split ? bz % g.splits : 0;
  tc.m0 = by * BM;
  tc.n0 = bx * BN;
  tc.A = g.A + (long long)tc.b * g.a_bs;
  tc.B = g.B + (long long)tc.b * g.b_bs;
  tc.K = g.K;
  tc.C = g.C + (long long)tc.b * g.c_bs;
  tc.c_bs = g.c_bs;
  tc.ldc = g.ldc;
  if (split) {
    const int k_lo = ks * g.kc;
    tc.A += (long long)k_lo * g.a_sk;
    tc.B += (long long)k_lo * g.b_sk;
    tc.K = (g.K - k_lo) < g.kc ? (g.K - k_lo) : g.kc;
    tc.C = g.split_buf + ((long long)ks * g.batch + tc.b) * (long long)g.M * g.N;
    tc.ldc = g.N;
  }
  tc.tile_linear = (long long)by * tiles_n + bx;
  return tc;
}

template <int EPI>
__global__ void __launch_bounds__(WS_THREADS, 1) tc_gemm_ws_kernel(GemmArgs g, int a_vec, int b_vec, int tiles_n, int tiles_m,
                                                                   long long total_tiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + WS_BAR_OFF);          // [WS_STAGES]
  uint64_t* empty = full + WS_STAGES;                                       // [WS_STAGES]
  uint64_t* tmem_full = empty + WS_STAGES;                                  // [2]
  uint64_t* tmem_empty = tmem_full + 2;                                     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* red = reinterpret_cast<float*>(tmem_slot + 4);                     // [2 * EPI_WARPS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < WS_STAGES; ++s) {
      mbar_init(&full[s], PRODUCER_WARPS);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PRODUCER_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PRODUCER_WARPS) {
    // ===================== producers =====================
    uint32_t it = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileCoord tc = decode_tile(g, t, tiles_n, tiles_m);
      TileLoader la{tc.A, g.a_sm, g.a_sk, g.M, tc.K, tc.m0, a_vec != 0};
      TileLoader lb{tc.B, g.b_sn, g.b_sk, g.N, tc.K, tc.n0, b_vec != 0};
      const int nk = (tc.K + BK - 1) / BK;
      float4 ra[2], rb[2];
      la.fetch(0, tid, ra);
      lb.fetch(0, tid, rb);
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const uint32_t s = it % WS_STAGES, ph = (it / WS_STAGES) & 1u;
        unsigned char* st = smem + s * STAGE_BYTES;
        mbar_wait(&empty[s], ph ^ 1u);
        stash(st, st + TILE_BYTES, tid, ra);
        stash(st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, tid, rb);
        if (kt + 1 < nk) {
          la.fetch((kt + 1) * BK, tid, ra);
          lb.fetch((kt + 1) * BK, tid, rb);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
  } else if (warp == PRODUCER_WARPS) {
    // ===================== MMA issuer =====================
    uint32_t it = 0, tcount = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tcount) {
      const TileCoord tc = decode_tile(g, t, tiles_n, tiles_m);
      const int nk = (tc.K + BK - 1) / BK;
      const uint32_t a = tcount & 1u;
      mbar_wait(&tmem_empty[a], ((tcount >> 1) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_main = tmem_base + a * 256u, d_corr = d_main + (uint32_t)BN;
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const uint32_t s = it % WS_STAGES, ph = (it / WS_STAGES) & 1u;
        mbar_wait(&full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const int steps = (tc.K - kt * BK) > 8 ? 2 : 1;
          for (int ks = 0; ks < steps; ++ks) {
            const uint32_t koff = (uint32_t)ks * 2u * LBO;
            const uint64_t a_hi = make_desc(sa + koff), a_lo = make_desc(sa + TILE_BYTES + koff);
            const uint64_t b_hi = make_desc(sa + 2 * TILE_BYTES + koff), b_lo = make_desc(sa + 3 * TILE_BYTES + koff);
            const uint32_t acc = (kt > 0 || ks > 0) ? 1u : 0u;
            mma_tf32(d_main, a_hi, b_hi, acc);
            mma_tf32(d_corr, a_lo, b_hi, acc);
            mma_tf32(d_corr, a_hi, b_lo, 1u);
          }
          mma_commit(&empty[s]);                       // stage reusable once these MMAs have read it
          if (kt == nk - 1) mma_commit(&tmem_full[a]); // accumulators of this tile complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue =====================
    const int e = warp - PRODUCER_WARPS - 1;          // 0..7
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int half = e >> 2;                          // which 64 accumulator columns this warp drains
    float* tile = reinterpret_cast<float*>(smem + WS_EPI_OFF);
    uint32_t tcount = 0;
    for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tcount) {
      const TileCoord tc = decode_tile(g, t, tiles_n, tiles_m);
      const uint32_t a = tcount & 1u;
      mbar_wait(&tmem_full[a], (tcount >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // phase 1: TMEM -> staging tile (thread = accumulator row q*32+lane, 64 columns)
      {
        float* trow = tile + (q * 32 + lane) * TILE_LD + half * (BN / 2);
        const uint32_t tbase = tmem_base + a * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * (BN / 2));
#pragma unroll 2
        for (int cc = 0; cc < BN / 2; cc += 8) {
          uint32_t r[8], rc[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                       : "r"(tbase + (uint32_t)cc));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(rc[0]), "=r"(rc[1]), "=r"(rc[2]), "=r"(rc[3]), "=r"(rc[4]), "=r"(rc[5]), "=r"(rc[6]), "=r"(rc[7])
                       : "r"(tbase + (uint32_t)(BN + cc)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) + __uint_as_float(rc[j]);
          *reinterpret_cast<float4*>(trow + cc) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(trow + cc + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      epi_bar();                                       // tile staged, every TMEM read of this set retired
      if (e == 0 && lane == 0) mbar_arrive(&tmem_empty[a]);
      // phase 2: row-wise coalesced epilogue
      const float bias0 = (EPI == EPI_HEAD) ? __ldg(g.bias + (long long)tc.b * g.bias_bs) : 0.0f;
      const float* auxb = (EPI == EPI_DACT || EPI == EPI_HEAD) ? g.aux + (long long)tc.b * g.aux_bs : nullptr;
      const float* biasb = (EPI == EPI_BIAS_ACT) ? g.bias + (long long)tc.b * g.bias_bs : nullptr;
      const bool vec = rows_vec_ok(tc.C, tc.c_bs, tc.ldc, g.N) && (tc.n0 % 4 == 0);
      const bool vec_aux = auxb != nullptr && rows_vec_ok(g.aux, g.aux_bs, g.ld_aux, g.N);
      float ll_acc = 0.0f, g_acc = 0.0f;
#pragma unroll 2
      for (int r = e; r < BM; r += EPI_WARPS) {
        const int m = tc.m0 + r;
        if (m >= g.M) break;
        epilogue_row<EPI>(g, tile + r * TILE_LD, tc.C + (long long)m * tc.ldc + tc.n0,
                          auxb ? auxb + (long long)m * g.ld_aux + tc.n0 : nullptr, biasb, tc.n0, lane, vec, vec_aux, bias0, ll_acc,
                          g_acc);
      }
      if (EPI == EPI_HEAD) {
        ll_acc = warp_sum(ll_acc);
        g_acc = warp_sum(g_acc);
        if (lane == 0) { red[e] = ll_acc; red[EPI_WARPS + e] = g_acc; }
      }
      epi_bar();                                       // staging tile free for the next tile; partial sums visible
      if (EPI == EPI_HEAD && e == 0 && lane == 0) {
        float s0 = 0.0f, s1 = 0.0f;
        for (int w = 0; w < EPI_WARPS; ++w) { s0 += red[w]; s1 += red[EPI_WARPS + w]; }
        const long long tiles = (long long)tiles_n * tiles_m;
        g.part_ll[(long long)tc.b * tiles + tc.tile_linear] = s0;
        g.part_g[(long long)tc.b * tiles + tc.tile_linear] = s1;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == PRODUCER_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace tcws

// C[b, m, n] = sum_s split_buf[s, b, m, n] in fixed order (fp32 round-to-nearest adds)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ buf, int splits, int batch, int M, int N,
                                                            float* __restrict__ C, long long c_bs, long long ldc) {
  const long long per = (long long)M * N, total = (long long)batch * per;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int k = 0; k < splits; ++k) s += buf[(long long)k * total + t];
    const long long b = t / per, r = t % per;
    C[b * c_bs + (r / N) * ldc + (r % N)] = s;
  }
}

constexpr int kSplitKChunk = 1024;      // K slice per CTA once K exceeds kSplitKThreshold
constexpr int kSplitKThreshold = 2048;

// floats of scratch a split-K GEMM of this shape needs (0 if it will not be split)
inline long long splitk_scratch_floats(int M, int N, int K, int batch) {
  if (K <= kSplitKThreshold) return 0;
  const int splits = (K + kSplitKChunk - 1) / kSplitKChunk;
  return (long long)splits * batch * M * N;
}

// shapes the tensor-core kernel is used for; everything else stays on the FP32-SIMT kernel
inline bool tc_gemm_eligible(const GemmArgs& g) { return g.M >= 32 && g.K >= 16 && (g.N >= 16 || g.K >= 512); }

template <int EPI>
static int launch_tc_gemm(GemmArgs g, int batch, cudaStream_t st, float* scratch = nullptr) {
  auto aligned16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const int a_vec = g.a_sk == 1 && g.a_sm % 4 == 0 && g.a_bs % 4 == 0 && g.K % 4 == 0 && aligned16(g.A);
  const int b_vec = g.b_sk == 1 && g.b_sn % 4 == 0 && g.b_bs % 4 == 0 && g.K % 4 == 0 && aligned16(g.B);
  auto k = tc::tc_gemm_kernel<EPI>;
  static bool configured = false;
  if (!configured) {
    VIHMC_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    configured = true;
  }
  g.splits = 1;
  g.batch = batch;
  if (EPI == EPI_STORE && scratch != nullptr && g.K > kSplitKThreshold) {
    // long reductions (dW = dZ^T X over thousands of rows): bound the length of one TMEM accumulation chain
    // (its fp32 accumulate truncates) and sum the slices with round-to-nearest adds
    g.kc = kSplitKChunk;
    g.splits = (g.K + g.kc - 1) / g.kc;
    g.split_buf = scratch;
    if ((long long)batch * g.splits > 65535) return fail(VIHMC_ERR_UNSUPPORTED, "gemm: batch * splits > 65535");
  }
  // The persistent warp-specialised kernel (tcws) is opt-in: measured 21.4 ms vs 18.6 ms per 64-chain DeepONet
  // gradient batch -- with K = 100 the epilogue is as heavy as the main loop, and 8 of 17 warps doing it lose to
  // 16 warps (2 CTAs) that all take part in every phase.
  static const bool simple = []() {
    const char* e = getenv("VIHMC_TC_PERSISTENT");
    return !(e != nullptr && e[0] == '1');
  }();
  const int tiles_n = (g.N + tc::BN - 1) / tc::BN, tiles_m = (g.M + tc::BM - 1) / tc::BM;
  if (simple) {
    dim3 grid(tiles_n, tiles_m, batch * g.splits);
    k<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(g, a_vec, b_vec);
    VIHMC_LAUNCH_OK("tc_gemm_kernel");
  } else {
    auto kw = tcws::tc_gemm_ws_kernel<EPI>;
    static bool configured_ws = false;
    if (!configured_ws) {
      VIHMC_CUDA_OK(cudaFuncSetAttribute(kw, cudaFuncAttributeMaxDynamicSharedMemorySize, tcws::WS_SMEM_BYTES));
      configured_ws = true;
    }
    const long long total = (long long)tiles_n * tiles_m * batch * g.splits;
    static const int num_sms = []() {
      int dev = 0, n = 148;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      return n;
    }();
    const int grid = (int)(total < num_sms ? total : num_sms);   // persistent: one CTA per SM
    kw<<<grid, tcws::WS_THREADS, tcws::WS_SMEM_BYTES, st>>>(g, a_vec, b_vec, tiles_n, tiles_m, total);
    VIHMC_LAUNCH_OK("tc_gemm_ws_kernel");
  }
  if (g.splits > 1) {
    const long long total = (long long)batch * g.M * g.N;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    splitk_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(scratch, g.splits, batch, g.M, g.N, g.C, g.c_bs, g.ldc);
    VIHMC_LAUNCH_OK("splitk_reduce_kernel");
  }
  return VIHMC_OK;
}

}  // namespace vihmc
