// ref_cuda_driver.cu -- C entry points around the REFERENCE's own CSR5_cuda backend.
//
// TEST INFRASTRUCTURE ONLY (see oracle/csr5_oracle.c header).  Contains no SpMV or format code of its
// own: it #includes the reference's header-only implementation (anonymouslib_cuda.h, from the
// compat-patched scratch copy made by oracle/build_ref_cuda.sh) and drives it the way the reference's
// call site does (CSR5_cuda/main.cu:17-117): H2D copies, inputCSR, setX, setSigma, asCSR5, ONE
// spmv(alpha, d_y) on a zeroed d_y, D2H of y.  In addition it reads back the handle's CSR5 arrays
// (private members, anonymouslib_cuda.h:25-52) so that the golden fixtures hold tile_ptr / tile_desc /
// the empty-row offset table / the transposed col and val exactly as the reference built them.
// warmup() is not called: it is an int function without a return statement (anonymouslib_cuda.h:56-59).
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <unistd.h>

using namespace std;  // the reference headers use cout/endl unqualified (main.cu provides this)

#define private public  // read-only access to the handle's CSR5 arrays
#include "anonymouslib_cuda.h"
#undef private

namespace {
struct quiet_stdout {
    std::streambuf *old_buf;
    std::ostringstream sink;
    quiet_stdout() { old_buf = std::cout.rdbuf(sink.rdbuf()); }
    ~quiet_stdout() { std::cout.rdbuf(old_buf); }
};

// scalars[]: sigma, bit_y_offset, bit_scansum_offset, num_packet, p, num_offsets, tail_partition_start
template <typename VT>
int run(int m, int n, int nnz, const int *row_ptr, const int *col, const VT *val, const VT *x, VT *y,
        int sigma, int ncalls, int *scalars, unsigned *tile_ptr, size_t tile_ptr_cap, unsigned *desc, size_t desc_cap,
        int *desc_off_ptr, size_t dop_cap, int *desc_off, size_t doff_cap, int *col5, VT *val5)
{
    quiet_stdout q;
    int *d_rp, *d_col;
    VT *d_val, *d_x, *d_y;
    checkCudaErrors(cudaMalloc(&d_rp, (size_t)(m + 1) * sizeof(int)));
    checkCudaErrors(cudaMalloc(&d_col, (size_t)nnz * sizeof(int)));
    checkCudaErrors(cudaMalloc(&d_val, (size_t)nnz * sizeof(VT)));
    checkCudaErrors(cudaMalloc(&d_x, (size_t)n * sizeof(VT)));
    checkCudaErrors(cudaMalloc(&d_y, (size_t)m * sizeof(VT)));
    checkCudaErrors(cudaMemcpy(d_rp, row_ptr, (size_t)(m + 1) * sizeof(int), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_col, col, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_val, val, (size_t)nnz * sizeof(VT), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_x, x, (size_t)n * sizeof(VT), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemset(d_y, 0, (size_t)m * sizeof(VT)));

    anonymouslibHandle<int, unsigned int, VT> A(m, n);
    int err = A.inputCSR(nnz, d_rp, d_col, d_val);
    if (!err) err = A.setX(d_x);
    A.setSigma(sigma);
    if (!err) err = A.asCSR5();
    checkCudaErrors(cudaDeviceSynchronize());
    if (!err) {
        scalars[0] = A._csr5_sigma; scalars[1] = A._bit_y_offset; scalars[2] = A._bit_scansum_offset;
        scalars[3] = A._num_packet; scalars[4] = A._p; scalars[5] = A._num_offsets;
        scalars[6] = A._tail_partition_start;
        const size_t p = (size_t)A._p, nd = p * 32 * A._num_packet;
        if (tile_ptr && tile_ptr_cap >= p + 1)
            checkCudaErrors(cudaMemcpy(tile_ptr, A._csr5_partition_pointer, (p + 1) * 4, cudaMemcpyDeviceToHost));
        if (desc && desc_cap >= nd)
            checkCudaErrors(cudaMemcpy(desc, A._csr5_partition_descriptor, nd * 4, cudaMemcpyDeviceToHost));
        if (desc_off_ptr && dop_cap >= p + 1)
            checkCudaErrors(cudaMemcpy(desc_off_ptr, A._csr5_partition_descriptor_offset_pointer, (p + 1) * 4,
                                       cudaMemcpyDeviceToHost));
        if (desc_off && A._num_offsets > 0 && doff_cap >= (size_t)A._num_offsets)
            checkCudaErrors(cudaMemcpy(desc_off, A._csr5_partition_descriptor_offset, (size_t)A._num_offsets * 4,
                                       cudaMemcpyDeviceToHost));
        if (col5) checkCudaErrors(cudaMemcpy(col5, d_col, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost));
        if (val5) checkCudaErrors(cudaMemcpy(val5, d_val, (size_t)nnz * sizeof(VT), cudaMemcpyDeviceToHost));
        for (int i = 0; i < ncalls && !err; i++) err = A.spmv((VT)1.0, d_y);  // no re-zeroing, as main.cu:84-99
        checkCudaErrors(cudaDeviceSynchronize());
        checkCudaErrors(cudaMemcpy(y, d_y, (size_t)m * sizeof(VT), cudaMemcpyDeviceToHost));
    }
    A.destroy();
    checkCudaErrors(cudaDeviceSynchronize());
    cudaFree(d_rp); cudaFree(d_col); cudaFree(d_val); cudaFree(d_x); cudaFree(d_y);
    return err;
}

// Timing protocol of main.cu:79-106 on device-resident data: 1 call, `warmup` calls, `runs` calls
// between two events.  (y drifts across calls in the reference -- SURVEY.md s0-2 -- which does not
// change the time.)
template <typename VT>
int bench(int m, int n, int nnz, const int *row_ptr, const int *col, const VT *val, const VT *x, int sigma,
          int warmup, int runs, double *ms_per_spmv, double *convert_ms)
{
    quiet_stdout q;
    int *d_rp, *d_col;
    VT *d_val, *d_x, *d_y;
    checkCudaErrors(cudaMalloc(&d_rp, (size_t)(m + 1) * sizeof(int)));
    checkCudaErrors(cudaMalloc(&d_col, (size_t)nnz * sizeof(int)));
    checkCudaErrors(cudaMalloc(&d_val, (size_t)nnz * sizeof(VT)));
    checkCudaErrors(cudaMalloc(&d_x, (size_t)n * sizeof(VT)));
    checkCudaErrors(cudaMalloc(&d_y, (size_t)m * sizeof(VT)));
    checkCudaErrors(cudaMemcpy(d_rp, row_ptr, (size_t)(m + 1) * sizeof(int), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_col, col, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_val, val, (size_t)nnz * sizeof(VT), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemcpy(d_x, x, (size_t)n * sizeof(VT), cudaMemcpyHostToDevice));
    checkCudaErrors(cudaMemset(d_y, 0, (size_t)m * sizeof(VT)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    anonymouslibHandle<int, unsigned int, VT> A(m, n);
    int err = A.inputCSR(nnz, d_rp, d_col, d_val);
    if (!err) err = A.setX(d_x);
    A.setSigma(sigma);
    checkCudaErrors(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    if (!err) err = A.asCSR5();
    cudaEventRecord(e1);
    checkCudaErrors(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    *convert_ms = ms;
    for (int i = 0; i < warmup + 1 && !err; i++) err = A.spmv((VT)1.0, d_y);
    checkCudaErrors(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < runs && !err; i++) err = A.spmv((VT)1.0, d_y);
    cudaEventRecord(e1);
    checkCudaErrors(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_per_spmv = runs > 0 ? ms / runs : 0.0;
    A.destroy();
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_rp); cudaFree(d_col); cudaFree(d_val); cudaFree(d_x); cudaFree(d_y);
    return err;
}
}  // namespace

extern "C" {

int ref_cuda_spmv_f64(int m, int n, int nnz, const int *row_ptr, const int *col, const double *val,
                      const double *x, double *y, int sigma, int ncalls, int *scalars, unsigned *tile_ptr,
                      size_t tile_ptr_cap, unsigned *desc, size_t desc_cap, int *desc_off_ptr, size_t dop_cap,
                      int *desc_off, size_t doff_cap, int *col5, double *val5)
{
    return run<double>(m, n, nnz, row_ptr, col, val, x, y, sigma, ncalls, scalars, tile_ptr, tile_ptr_cap, desc,
                       desc_cap, desc_off_ptr, dop_cap, desc_off, doff_cap, col5, val5);
}

int ref_cuda_spmv_f32(int m, int n, int nnz, const int *row_ptr, const int *col, const float *val,
                      const float *x, float *y, int sigma, int ncalls, int *scalars, unsigned *tile_ptr,
                      size_t tile_ptr_cap, unsigned *desc, size_t desc_cap, int *desc_off_ptr, size_t dop_cap,
                      int *desc_off, size_t doff_cap, int *col5, float *val5)
{
    return run<float>(m, n, nnz, row_ptr, col, val, x, y, sigma, ncalls, scalars, tile_ptr, tile_ptr_cap, desc,
                      desc_cap, desc_off_ptr, dop_cap, desc_off, doff_cap, col5, val5);
}

int ref_cuda_bench_f64(int m, int n, int nnz, const int *row_ptr, const int *col, const double *val,
                       const double *x, int sigma, int warmup, int runs, double *ms_per_spmv, double *convert_ms)
{
    return bench<double>(m, n, nnz, row_ptr, col, val, x, sigma, warmup, runs, ms_per_spmv, convert_ms);
}

int ref_cuda_bench_f32(int m, int n, int nnz, const int *row_ptr, const int *col, const float *val,
                       const float *x, int sigma, int warmup, int runs, double *ms_per_spmv, double *convert_ms)
{
    return bench<float>(m, n, nnz, row_ptr, col, val, x, sigma, warmup, runs, ms_per_spmv, convert_ms);
}

}  // extern "C"
