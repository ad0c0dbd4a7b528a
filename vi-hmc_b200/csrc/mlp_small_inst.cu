// One translation unit per padded hidden width (compiled with -DVIHMC_W=10|16|32) so the heavy,
// fully unrolled kernels build in parallel.
#include "mlp_small.cuh"

#define VIHMC_CAT2(a, b) a##b
#define VIHMC_CAT(a, b) VIHMC_CAT2(a, b)

namespace vihmc {
int VIHMC_CAT(mlp_small_launch_w, VIHMC_W)(SmallOp op, const SmallParams& P, const SmallLaunch& a, cudaStream_t st) {
  return launch_small_w<VIHMC_W>(op, P, a, st);
}
}  // namespace vihmc
