// FP64 instantiations of the CSR5 SpMV kernels (sigma 4..32, direct-load and TMA-staged).
#include "csr5_spmv.cuh"

namespace csr5 {
cudaError_t launch_spmv_f64(const Plan &pl, const SpmvTuning &tn, double alpha, double *y, const ShardCtx *sh,
                            cudaStream_t stream, int *used, int *launches)
{
    return launch_spmv_t<double>(pl, tn, alpha, y, sh, stream, used, launches);
}
}  // namespace csr5
