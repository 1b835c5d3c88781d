// csr5_internal.h -- state and launcher prototypes shared by the translation units of
// libcsr5_b200.so.  Nothing here is part of the ABI (see include/csr5_b200.h).
#ifndef CSR5_INTERNAL_H
#define CSR5_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/csr5_b200.h"

namespace csr5 {

constexpr int OMEGA = CSR5B200_OMEGA;
constexpr uint32_t MSB = 0x80000000u;
constexpr uint32_t ROW_MASK = 0x7FFFFFFFu;
constexpr int SIGMA_MIN = 4;   // csr5_spmv_cuda.h:448-540: the reference instantiates sigma 4..32
constexpr int SIGMA_MAX = 32;

// Mirrors the private section of the reference handle (anonymouslib_cuda.h:25-52).
struct Plan {
    int m = 0, n = 0, nnz = 0;
    int value_bytes = 8;
    int sigma = 0;          // 0 = never set (the reference leaves it uninitialised)
    int bit_y = 0, bit_ss = 0, num_packet = 0;
    int p = 0;
    int tail_start = 0;
    int num_offsets = 0;
    int needs_zero_fill = 0;

    const int *row_ptr = nullptr;   // borrowed
    int *col = nullptr;             // borrowed, permuted in place while in CSR5 format
    void *val = nullptr;            // borrowed, permuted in place while in CSR5 format
    const void *x = nullptr;        // borrowed

    uint32_t *tile_ptr = nullptr;   // owned, p + 1
    uint32_t *desc = nullptr;       // owned, p * 32 * num_packet
    int *desc_off_ptr = nullptr;    // owned, p + 1
    int *desc_off = nullptr;        // owned, num_offsets
    void *calibrator = nullptr;     // owned, p values
    int *dev_flags = nullptr;       // owned, small scratch: [0] any dirty tile before the tail
};

struct SpmvTuning {
    int kernel = 0;        // 0 auto, 1 direct-load, 2 TMA-staged
    int tma_stages = 0;
    int tma_warps = 0;
    int ctas_per_sm = 0;
    int num_sms = 148;
    int direct_wpb = 0;    // tuning: warps per CTA of the direct kernel (0 = default 4)
    int direct_nch = 0;    // tuning: register chunks per tile (0 = default rule)
    cudaEvent_t ev_begin = nullptr;  // optional: recorded right before / after the main SpMV kernel
    cudaEvent_t ev_end = nullptr;
};

// ---- format conversion (csr5_format.cu) ------------------------------------------------------
// tile_ptr / descriptor / segment counts; leaves the exclusive scan and offset table to the next
// two calls.  All asynchronous on `stream`.
cudaError_t launch_tile_ptr(const Plan &pl, cudaStream_t stream);
cudaError_t launch_tile_desc(const Plan &pl, cudaStream_t stream);
cudaError_t launch_scan_offsets(const Plan &pl, void *scratch, size_t scratch_bytes, cudaStream_t stream);
size_t scan_scratch_bytes(int p);
cudaError_t launch_desc_offset(const Plan &pl, cudaStream_t stream);
// in-place omega x sigma tile transpose of col and val; r2c = CSR -> CSR5.
cudaError_t launch_transpose(const Plan &pl, bool r2c, cudaStream_t stream);
cudaError_t launch_warmup(cudaStream_t stream);

// ---- SpMV (csr5_spmv_f64.cu / csr5_spmv_f32.cu via csr5_spmv.cuh) -----------------------------
// Enqueues [clear y] + compute(+tail) + calibrate.  Returns the kernel variant used in *used.
// n_dst == 0: y is the (local) result vector.  n_dst > 0: y is ignored and every value is stored to
// y_dst[0..n_dst) instead (sharded mode: this rank's segment inside each peer's concatenated y).
cudaError_t launch_spmv_f64(const Plan &pl, const SpmvTuning &tn, double alpha, double *y, int n_dst,
                            void *const *y_dst, cudaStream_t stream, int *used, int *launches);
cudaError_t launch_spmv_f32(const Plan &pl, const SpmvTuning &tn, float alpha, float *y, int n_dst,
                            void *const *y_dst, cudaStream_t stream, int *used, int *launches);

}  // namespace csr5

#endif
