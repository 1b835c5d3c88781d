// FP64 instantiations of the CSR5 SpMV kernels (sigma 4..32: direct-load, TMA-staged, hot-column).
#include "csr5_spmv.cuh"

namespace csr5 {
cudaError_t launch_spmv_part_f64(const Plan &pl, const SpmvTuning &tn, double alpha, double beta, double *y,
                                 const ShardCtx *sh, const SpmvCall &call, cudaStream_t stream, int *used, int *launches)
{
    return launch_spmv_part_t<double>(pl, tn, alpha, beta, y, sh, call, stream, used, launches);
}
cudaError_t launch_push_rows_f64(const void *y_local, void *const *dst, int n_dst, int multicast, long long rows,
                                 int grid, int threads, cudaStream_t stream)
{
    return launch_push_t<double>(static_cast<const double *>(y_local), reinterpret_cast<double *const *>(dst), n_dst, multicast,
                             rows, grid, threads, stream);
}
cudaError_t launch_push_row_list_f64(const void *y_local, void *const *dst, int n_dst, int multicast, const int *rows,
                                     int n, cudaStream_t stream)
{
    return launch_push_row_list_t<double>(static_cast<const double *>(y_local), reinterpret_cast<double *const *>(dst), n_dst,
                                      multicast, rows, n, stream);
}
}  // namespace csr5
