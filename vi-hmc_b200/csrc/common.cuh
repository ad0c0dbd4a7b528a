// Shared device/host helpers for libvihmc (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vihmc.h"

namespace vihmc {

// ---------------------------------------------------------------------------------------------
// error plumbing (thread-local message behind vihmc_last_error)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define VIHMC_CUDA_OK(expr)                                                                         \
  do {                                                                                              \
    cudaError_t err__ = (expr);                                                                     \
    if (err__ != cudaSuccess)                                                                       \
      return ::vihmc::fail(VIHMC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                           __FILE__, __LINE__);                                                     \
  } while (0)

#define VIHMC_LAUNCH_OK(what)                                                                       \
  do {                                                                                              \
    cudaError_t err__ = cudaGetLastError();                                                         \
    if (err__ != cudaSuccess)                                                                       \
      return ::vihmc::fail(VIHMC_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(err__)); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  counter = (c0,c1,c2,c3), key = (k0,k1).
// Engine convention: key = 64-bit seed; counter = (global chain id, iteration, block, stream).
// ---------------------------------------------------------------------------------------------
enum : uint32_t { STREAM_MOMENTUM = 0, STREAM_UNIFORM = 1, STREAM_VI_REDRAW = 2, STREAM_INIT = 3 };

struct u32x4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c0), hi1 = __umulhi(M1, c2);
#else
    uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c0) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c2) >> 32);
#endif
    uint32_t lo0 = M0 * c0, lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return u32x4{c0, c1, c2, c3};
}

// 24-bit uniform strictly inside (0,1): exactly representable in fp32, log() always finite.
__host__ __device__ __forceinline__ float u32_to_unit(uint32_t v) {
  return ((float)(v >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// Box-Muller on two 32-bit words -> two N(0,1).  Accurate (non-fast-math) logf/sincospif.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  float u1 = u32_to_unit(a), u2 = u32_to_unit(b);
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// four normals for coordinates 4*block .. 4*block+3 of (chain, iteration, stream)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t chain, uint32_t iteration, uint32_t block,
                                                 uint32_t stream) {
  // 64-bit chain ids are folded: low word in c0, high word xored into the stream word's upper bits.
  u32x4 r = philox4x32_10((uint32_t)chain, iteration, block, stream ^ ((uint32_t)(chain >> 32) << 8),
                          (uint32_t)seed, (uint32_t)(seed >> 32));
  float4 z;
  box_muller(r.x, r.y, z.x, z.y);
  box_muller(r.z, r.w, z.z, z.w);
  return z;
}

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t chain, uint32_t iteration) {
  u32x4 r = philox4x32_10((uint32_t)chain, iteration, 0u, STREAM_UNIFORM ^ ((uint32_t)(chain >> 32) << 8),
                          (uint32_t)seed, (uint32_t)(seed >> 32));
  return u32_to_unit(r.x);
}

// ---------------------------------------------------------------------------------------------
// warp reductions: fixed butterfly order => bit-reproducible, identical on every lane
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// activations (my_make_func.py:36-43): value and derivative w.r.t. pre-activation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(int act, float z, float& dact) {
  if (act == VIHMC_ACT_TANH) {
    float h = tanhf(z);
    dact = 1.0f - h * h;
    return h;
  } else if (act == VIHMC_ACT_RELU) {
    dact = z > 0.0f ? 1.0f : 0.0f;
    return z > 0.0f ? z : 0.0f;
  } else {
    float s, c;
    sincosf(z, &s, &c);
    dact = c;
    return s;
  }
}

// tanh(x) = 1 - 2 / (2^(x * 2 log2 e) + 1) on ex2.approx / rcp.approx: 5 instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA), no
// branch, no sign handling: e = +inf for large x gives rcp = 0 and tanh = 1, e = 0 (flushed) for large -x gives 1 - 2 = -1.
// Absolute error <= ~1.2e-7 (one ulp at 1.0) over the whole range -- the same as tanhf's large-|x| branch; near 0 the
// RELATIVE error grows like 6e-8/|x|, which is harmless here because activations only ever enter sums against O(1) terms
// (checked by the rtol-1e-5 parity tests against the reference).  The result is odd in x only to that accuracy.
__device__ __forceinline__ float tanh_sel(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

// packed FP32 pairs (Blackwell FFMA2: one instruction, two IEEE fma.rn results)
__device__ __forceinline__ unsigned long long pack_f2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma_f2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// two tanh_sel at once: the three FP32 steps on the packed pipe (7 instructions per pair instead of 10); every step is the
// same IEEE operation as in tanh_sel (x * c = fma(x, c, 0), e + 1 = fma(e, 1, 1)), so the results are bit-identical to it
__device__ __forceinline__ unsigned long long tanh_sel2(unsigned long long x) {
  const unsigned long long c = pack_f2(2.885390081777927f, 2.885390081777927f), one = pack_f2(1.0f, 1.0f), m2 = pack_f2(-2.0f, -2.0f);
  float a0, a1, e0, e1, r0, r1;
  unpack_f2(fma_f2(x, c, 0ull), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  unpack_f2(fma_f2(pack_f2(e0, e1), one, one), a0, a1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(a0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(a1));
  return fma_f2(pack_f2(r0, r1), m2, one);
}

// Gaussian likelihood pieces (main_VI_HMC.py:132-136; GaussianNLLLoss clamps var at 1e-6, full=False)
struct Likelihood {
  float ll_const;   // per-output additive constant: NLL: -0.5 log v ; regression: 0
  float half_prec;  // NLL: 0.5 / v ; regression: 0.5 tau
  float prec;       // NLL: 1 / v   ; regression: tau          (d loglik / d o = -prec (o - y))
};

__host__ __device__ __forceinline__ Likelihood make_likelihood(int loss, float tau_out) {
  Likelihood l;
  if (loss == VIHMC_LOSS_NLL) {
    float v = tau_out < 1e-6f ? 1e-6f : tau_out;
    l.ll_const = -0.5f * logf(v);
    l.half_prec = 0.5f / v;
    l.prec = 1.0f / v;
  } else {
    l.ll_const = 0.0f;
    l.half_prec = 0.5f * tau_out;
    l.prec = tau_out;
  }
  return l;
}

// hamiltorch writes `p += c * g` as two separately rounded torch ops (mul kernel, add kernel):
// mirror that rounding instead of letting nvcc contract to an FMA.
__device__ __forceinline__ float axpy_unfused(float a, float x, float y) { return __fadd_rn(y, __fmul_rn(a, x)); }

}  // namespace vihmc
