// anonymouslib_cuda.h -- drop-in replacement for the reference's CSR5_cuda/anonymouslib_cuda.h.
//
// Same class template, same public methods, same signatures and return codes as the reference's
// anonymouslibHandle<IT, UIT, VT> (CSR5_cuda/anonymouslib_cuda.h:11-24), so the reference's call
// site -- call_anonymouslib() in CSR5_cuda/main.cu:17-117 -- compiles UNCHANGED against this header
// and links against libcsr5_b200.so (C ABI in csr5_b200.h).  The class is a thin inline shim: no
// kernel lives here; every method forwards to one extern "C" entry point.
//
// Also provided, because the reference's main.cu expects them from the same include:
//   anonymouslib_timer            (detail/cuda/utils_cuda.h:6-23; here the events are destroyed)
//   getB<iT, vT>, getFLOP<iT>     (detail/utils.h:10-20; the reference's bandwidth/flop formulas)
//   ANONYMOUSLIB_* codes          (detail/common.h:13-22, detail/cuda/common_cuda.h:11-15)
//   checkCudaErrors               (CUDA-samples helper_cuda.h, which the reference does not vendor)
//
// Build:  nvcc -arch=sm_100a -I<this dir> main.cu -L<dir of libcsr5_b200.so> -lcsr5_b200
#ifndef ANONYMOUSLIB_CUDA_H
#define ANONYMOUSLIB_CUDA_H

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "csr5_b200.h"

using namespace std;  // the reference's detail/common.h:10 does this, and its main.cu relies on it

#define ANONYMOUSLIB_SUCCESS                 CSR5B200_SUCCESS
#define ANONYMOUSLIB_UNKOWN_FORMAT           CSR5B200_UNKNOWN_FORMAT
#define ANONYMOUSLIB_UNSUPPORTED_CSR5_OMEGA  CSR5B200_UNSUPPORTED_CSR5_OMEGA
#define ANONYMOUSLIB_CSR_TO_CSR5_FAILED      CSR5B200_CSR_TO_CSR5_FAILED
#define ANONYMOUSLIB_UNSUPPORTED_CSR_SPMV    CSR5B200_UNSUPPORTED_CSR_SPMV
#define ANONYMOUSLIB_UNSUPPORTED_VALUE_TYPE  CSR5B200_UNSUPPORTED_VALUE_TYPE

#define ANONYMOUSLIB_FORMAT_CSR   CSR5B200_FORMAT_CSR
#define ANONYMOUSLIB_FORMAT_CSR5  CSR5B200_FORMAT_CSR5
#define ANONYMOUSLIB_FORMAT_HYB5  2

#define ANONYMOUSLIB_CSR5_OMEGA        CSR5B200_OMEGA
#define ANONYMOUSLIB_THREAD_BUNCH      32
#define ANONYMOUSLIB_THREAD_GROUP      128
#define ANONYMOUSLIB_AUTO_TUNED_SIGMA  CSR5B200_AUTO_TUNED_SIGMA

#ifndef checkCudaErrors
#define checkCudaErrors(call)                                                                       \
    do {                                                                                            \
        cudaError_t anonymouslib_err_ = (call);                                                     \
        if (anonymouslib_err_ != cudaSuccess) {                                                     \
            fprintf(stderr, "CUDA error at %s:%d code=%d \"%s\"\n", __FILE__, __LINE__,            \
                    (int)anonymouslib_err_, cudaGetErrorString(anonymouslib_err_));                 \
            exit(EXIT_FAILURE);                                                                     \
        }                                                                                           \
    } while (0)
#endif

// Bytes / flops of one SpMV as the reference counts them (x charged once per non-zero).
template <typename iT, typename vT>
double getB(const iT m, const iT nnz)
{
    return (double)(((double)m + 1 + (double)nnz) * sizeof(iT) + (2.0 * (double)nnz + (double)m) * sizeof(vT));
}

template <typename iT>
double getFLOP(const iT nnz)
{
    return 2.0 * (double)nnz;
}

// start()/stop() pair on the legacy default stream; stop() returns milliseconds.
struct anonymouslib_timer {
    cudaEvent_t start_event = nullptr, stop_event = nullptr;

    void start()
    {
        release();
        cudaEventCreate(&start_event);
        cudaEventCreate(&stop_event);
        cudaDeviceSynchronize();
        cudaEventRecord(start_event, 0);
    }

    float stop()
    {
        float ms = 0.f;
        cudaEventRecord(stop_event, 0);
        cudaEventSynchronize(stop_event);
        cudaEventElapsedTime(&ms, start_event, stop_event);
        release();
        return ms;
    }

    void release()
    {
        if (start_event) cudaEventDestroy(start_event);
        if (stop_event) cudaEventDestroy(stop_event);
        start_event = stop_event = nullptr;
    }
};

template <class ANONYMOUSLIB_IT, class ANONYMOUSLIB_UIT, class ANONYMOUSLIB_VT>
class anonymouslibHandle
{
    static_assert(sizeof(ANONYMOUSLIB_IT) == 4 && sizeof(ANONYMOUSLIB_UIT) == 4,
                  "libcsr5_b200 is built for 32-bit indices, as the reference's only instantiation "
                  "anonymouslibHandle<int, unsigned int, VALUE_TYPE> (CSR5_cuda/main.cu:59)");
    static_assert(sizeof(ANONYMOUSLIB_VT) == 4 || sizeof(ANONYMOUSLIB_VT) == 8,
                  "VALUE_TYPE must be float or double (CSR5_cuda/Makefile:4)");

public:
    anonymouslibHandle(ANONYMOUSLIB_IT m, ANONYMOUSLIB_IT n)
    {
        _err = csr5b200_create((int)m, (int)n, (int)sizeof(ANONYMOUSLIB_VT), &_h);
    }
    // The reference handle has no destructor (destroy() is explicit); releasing the C object here is
    // safe because csr5b200_free() only restores/frees what destroy() has not already.
    ~anonymouslibHandle() { csr5b200_free(_h); }
    anonymouslibHandle(const anonymouslibHandle &) = delete;
    anonymouslibHandle &operator=(const anonymouslibHandle &) = delete;

    int warmup() { return _h ? csr5b200_warmup(_h) : _err; }
    int inputCSR(ANONYMOUSLIB_IT nnz, ANONYMOUSLIB_IT *csr_row_pointer, ANONYMOUSLIB_IT *csr_column_index,
                 ANONYMOUSLIB_VT *csr_value)
    {
        return _h ? csr5b200_input_csr(_h, (int)nnz, reinterpret_cast<int *>(csr_row_pointer),
                                       reinterpret_cast<int *>(csr_column_index), csr_value)
                  : _err;
    }
    int asCSR() { return _h ? csr5b200_as_csr(_h) : _err; }
    int asCSR5() { return _h ? csr5b200_as_csr5(_h) : _err; }
    int setX(ANONYMOUSLIB_VT *x) { return _h ? csr5b200_set_x(_h, x) : _err; }
    int spmv(const ANONYMOUSLIB_VT alpha, ANONYMOUSLIB_VT *y) { return _h ? csr5b200_spmv(_h, (double)alpha, y) : _err; }
    int destroy() { return _h ? csr5b200_destroy(_h) : _err; }
    void setSigma(int sigma)
    {
        if (_h) csr5b200_set_sigma(_h, sigma);
    }

    // not in the reference: access to the C handle for set_stream / set_option / get_info
    csr5b200_handle_t c_handle() const { return _h; }

private:
    csr5b200_handle_t _h = nullptr;
    int _err = ANONYMOUSLIB_SUCCESS;
};

#endif  // ANONYMOUSLIB_CUDA_H
