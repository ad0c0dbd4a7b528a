// Micro-probe (round 2): do warp shuffles share the shared-memory data pipe's wavefront budget with LDS?
// Three loops per warp: A = 8 LDS.128 per iteration, B = 32 SHFL.IDX per iteration, C = both.  If t(C) ~ max(t(A), t(B)) the two
// are separate resources; if t(C) ~ t(A) + t(B) a shuffle costs a wavefront of the same pipe.
// Build/run: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_shfl_lds tools/probe_shfl_lds.cu && /tmp/probe_shfl_lds
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void probe(float* out, int iters) {
  __shared__ float4 sm[32 * 8];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * 8; i += blockDim.x) sm[i] = make_float4(i, 1.0f, 2.0f, 3.0f);
  __syncthreads();
  float acc[8];
  float src[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { acc[k] = k; src[k] = lane + k; }
  const float4* base = sm + lane;   // every lane its own 16 bytes: an LDS.128 is 4 wavefronts
  for (int it = 0; it < iters; ++it) {
    if (MODE & 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 v = base[32 * ((k + it) & 7)];
        acc[k] += v.x + v.y + v.z + v.w;
      }
    }
    if (MODE & 2) {   // 32 independent shuffles (sources do not depend on earlier shuffles)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += __shfl_sync(0xffffffffu, src[k], (lane + r + it) & 31);
    }
  }
  float t = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
static float run(float* out, int warps_per_sm, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = 148 * warps_per_sm;
  probe<MODE><<<blocks, 32>>>(out, iters);
  cudaEventRecord(e0);
  probe<MODE><<<blocks, 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 64 * 32 * sizeof(float));
  const int iters = 20000;
  for (int w : {4, 7, 8, 16, 32}) {
    const float a = run<1>(out, w, iters), b = run<2>(out, w, iters), c = run<3>(out, w, iters);
    // per SM and iteration: w warps x (8 LDS.128 = 32 wavefronts | 32 SHFL)
    const double clk = 1.965e6;   // cycles per ms at 1965 MHz
    printf("{\"warps_per_sm\": %d, \"lds_ms\": %.3f, \"shfl_ms\": %.3f, \"both_ms\": %.3f, \"lds_wavefronts_per_clk_sm\": %.3f, "
           "\"shfl_per_clk_sm\": %.3f, \"both_over_sum\": %.3f, \"both_over_max\": %.3f}\n",
           w, a, b, c, w * 32.0 * iters / (a * clk), w * 32.0 * iters / (b * clk), c / (a + b), c / (a > b ? a : b));
  }
  return 0;
}
