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
// Structure of one CTA (256 threads, one 128 x 128 output tile, 64 KB smem, 256 TMEM columns => 2 CTAs / SM so one CTA's
// epilogue overlaps another's main loop):
//   * all 8 warps stream the A / B k-tiles (16 floats deep) from global memory through registers, split
//     them, and store hi/lo tiles in the UMMA canonical K-major no-swizzle layout (8x16B core matrices);
//     arbitrary element strides are supported, so transposed operands (dW = dZ^T X, dX = dZ W) and
//     rows that are not 16-byte aligned (101-wide branch input) need no extra copies -- this is also
//     why the operands are not staged with TMA: the split needs the values in registers anyway.
//   * two smem stages; thread 0 issues 3 MMAs per 8-deep k-step and commits to the stage's mbarrier;
//     the next tile's global loads are in flight while the tensor core works.
//   * epilogue: 8 warps read the accumulators with tcgen05.ld (warp w -> TMEM lanes 32*(w%4).., column
//     half w/4) into a shared-memory tile, then apply bias+act / act' / Gaussian residual and write C
//     row-wise so every warp instruction touches contiguous memory.
#pragma once
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

constexpr int BM = 128, BN = 128, BK = 16, THREADS = 256, STAGES = 2;
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
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
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
    const int s = kt & 1;
    unsigned char* st = smem + s * STAGE_BYTES;
    if (kt >= 2) mbar_wait(&mbar[s], (uint32_t)((kt >> 1) - 1) & 1u);   // the MMAs that read this stage are done
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
  mbar_wait(&mbar[(nk - 1) & 1], (uint32_t)((nk - 1) >> 1) & 1u);   // commits complete in order: everything is done
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
  float ll_acc = 0.0f, g_acc = 0.0f;
  float bias_r[4] = {0.f, 0.f, 0.f, 0.f};   // this lane's four columns: loaded once, not once per row
  if (EPI == EPI_BIAS_ACT) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n0 + lane + 32 * j < g.N) bias_r[j] = __ldg(biasb + n0 + lane + 32 * j);
  }
#pragma unroll 1
  for (int r = warp; r < BM; r += THREADS / 32) {
    const int m = m0 + r;
    if (m >= g.M) break;
    float* crow = Cb + (long long)m * g.ldc + n0;
    const float* arow = auxb ? auxb + (long long)m * g.ld_aux + n0 : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;           // 128 contiguous bytes per warp instruction
      const int n = n0 + c;
      if (n >= g.N) continue;
      float v = tile[r * TILE_LD + c];
      if (EPI == EPI_BIAS_ACT) {
        v += bias_r[j];
        if (g.act == VIHMC_ACT_TANH) v = tanh_sel(v);   // 8 instructions, abs. error ~1.2e-7 (see common.cuh)
        else if (g.act == VIHMC_ACT_RELU) v = v > 0.0f ? v : 0.0f;
      } else if (EPI == EPI_DACT) {
        const float a = __ldg(arow + c);
        v *= (g.act == VIHMC_ACT_TANH) ? (1.0f - a * a) : (a > 0.0f ? 1.0f : 0.0f);
      } else if (EPI == EPI_HEAD) {
        const float res = v + bias0 - __ldg(arow + c);
        ll_acc += g.ll_const - g.half_prec * res * res;
        v = -g.prec * res;
        g_acc += v;
      }
      crow[c] = v;
    }
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
  dim3 grid((g.N + tc::BN - 1) / tc::BN, (g.M + tc::BM - 1) / tc::BM, batch * g.splits);
  k<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(g, a_vec, b_vec);
  VIHMC_LAUNCH_OK("tc_gemm_kernel");
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
