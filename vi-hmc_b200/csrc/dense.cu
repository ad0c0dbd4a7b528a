// Dense path (DeepONet, wide MLP) -- placeholder until the GEMM kernels land.
#include "common.cuh"

namespace vihmc {
bool dense_supported(const vihmc_problem*) { return false; }
size_t dense_workspace_bytes(const vihmc_problem*, long long) { return 0; }
int dense_logp_grad(const vihmc_problem*, long long, const float*, float*, float*, void*, size_t, cudaStream_t) {
  return fail(VIHMC_ERR_UNSUPPORTED, "dense path not built");
}
int dense_predict(const vihmc_problem*, long long, const float*, float*, void*, size_t, cudaStream_t) {
  return fail(VIHMC_ERR_UNSUPPORTED, "dense path not built");
}
}  // namespace vihmc
