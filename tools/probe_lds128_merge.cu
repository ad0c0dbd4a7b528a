// Micro-probe (round 2): when does an LDS.128 of a full warp cost fewer than 4 shared-memory wavefronts?
// Every pattern maps lane -> 16-byte chunk index; the loop issues 8 independent LDS.128 per iteration.
// Build/run: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe_lds128_merge.bin tools/probe_lds128_merge.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int chunk_of(int pattern, int lane) {
  switch (pattern) {
    case 0: return lane & 7;                        // every quarter-warp the same 8 chunks, same order
    case 1: return (lane + (lane >> 3)) & 7;        // same 8 chunks, rotated per quarter-warp
    case 2: return (lane % 5) * 3;                 // 5 chunks 48 bytes apart (weight rows of the BNN kernel, lane = jp + 5 nq)
    case 3: return lane / 5;                        // 7 consecutive chunks, 5 lanes each (activation quads of the BNN kernel)
    case 4: return lane;                            // all distinct
    case 5: return lane & 15;                       // 16 distinct chunks (256 bytes), two lanes each
    case 6: return lane >> 2;                       // 8 chunks, 4 consecutive lanes each
    case 7: return 0;                               // full broadcast
    case 8: return (lane & 7) * 3;                  // 8 chunks 48 bytes apart, same in every quarter-warp
    case 9: return lane % 6;                        // 6 consecutive chunks, lane = nq + 6 jp
    default: return lane;
  }
}

__global__ void probe(float* out, int iters, int pattern) {
  __shared__ float4 sm[64 * 8];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) sm[i] = make_float4(i, 1.0f, 2.0f, 3.0f);
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = k;
  const float4* base = sm + chunk_of(pattern, lane);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = base[64 * ((k + it) & 7)];
      acc[k] += v.x + v.y + v.z + v.w;
    }
  }
  float t = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 64 * 32 * sizeof(float));
  const int iters = 20000, w = 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pattern = 0; pattern < 10; ++pattern) {
    probe<<<148 * w, 32>>>(out, iters, pattern);
    cudaEventRecord(e0);
    probe<<<148 * w, 32>>>(out, iters, pattern);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("{\"pattern\": %d, \"ms\": %.3f, \"cycles_per_lds128\": %.3f}\n", pattern, ms, ms * 1.965e6 / (w * 8.0 * iters));
  }
  return 0;
}
