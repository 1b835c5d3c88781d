// FP32 instantiations of the CSR5 SpMV kernels (sigma 4..32, direct-load and TMA-staged).
#include "csr5_spmv.cuh"

namespace csr5 {
cudaError_t launch_spmv_f32(const Plan &pl, const SpmvTuning &tn, float alpha, float *y, const ShardCtx *sh,
                            cudaStream_t stream, int *used, int *launches)
{
    return launch_spmv_t<float>(pl, tn, alpha, y, sh, stream, used, launches);
}
}  // namespace csr5
