// ref_avx2_driver.cpp -- C entry points around the REFERENCE's own CSR5_avx2 backend.
//
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/csr5_oracle.c header).  This file contains no
// SpMV code of its own for FP64: it #includes the reference's header-only implementation from
// where it lies (-I/root/reference/CSR5_avx2, anonymouslib_avx2.h) and drives it exactly as the
// reference's call site does (CSR5_avx2/main.cpp:18-89).  The result, oracle/_ref/libref_avx2.so,
// is git-ignored and travels to the GPU box with the snapshot.
//
// Constraints of the reference backend honoured here: sigma is the compile-time
// ANONYMOUSLIB_CSR5_SIGMA (16), omega 4, FP64 only; val must be 32-byte aligned; y must be zero
// on entry (empty rows are never written); asCSR5() permutes col/val in place, so the caller's
// arrays are copied into aligned scratch first.
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <sys/time.h>
#include <unistd.h>
#include <omp.h>

#include "anonymouslib_avx2.h"  // the reference, unmodified

namespace {

struct quiet_stdout {  // the reference prints conversion timings to stdout; keep test logs clean
    std::streambuf *old_buf;
    std::ostringstream sink;
    int saved_fd;
    quiet_stdout() {
        old_buf = std::cout.rdbuf(sink.rdbuf());
        fflush(stdout);
        saved_fd = dup(1);
        FILE *nul = fopen("/dev/null", "w");
        if (nul) { dup2(fileno(nul), 1); fclose(nul); }
    }
    ~quiet_stdout() {
        fflush(stdout);
        if (saved_fd >= 0) { dup2(saved_fd, 1); close(saved_fd); }
        std::cout.rdbuf(old_buf);
    }
};

double now_ms() {
    timeval tv;
    gettimeofday(&tv, nullptr);
    return tv.tv_sec * 1e3 + tv.tv_usec * 1e-3;
}

}  // namespace

extern "C" {

int ref_avx2_max_threads() { return omp_get_max_threads(); }
int ref_avx2_sigma() { return ANONYMOUSLIB_CSR5_SIGMA; }
int ref_avx2_omega() { return ANONYMOUSLIB_CSR5_OMEGA; }

// One checked SpMV: inputCSR -> setX -> setSigma(16) -> asCSR5 -> spmv(1.0, y) -> destroy,
// CSR5_avx2/main.cpp:28-57,83.  nthreads <= 0 keeps the OpenMP default.
int ref_avx2_spmv(int m, int n, int nnz, const int *row_ptr, const int *col, const double *val,
                  const double *x, double *y, int nthreads)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    int *rp = (int *)_mm_malloc((size_t)(m + 2) * sizeof(int), 64);  // +1: format_avx2.h:48-50 reads row_ptr[m+1]
    int *ci = (int *)_mm_malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int), 64);
    double *va = (double *)_mm_malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(double), 64);
    memcpy(rp, row_ptr, (size_t)(m + 1) * sizeof(int));
    rp[m + 1] = rp[m];
    memcpy(ci, col, (size_t)nnz * sizeof(int));
    memcpy(va, val, (size_t)nnz * sizeof(double));
    memset(y, 0, (size_t)m * sizeof(double));

    int err;
    {
        quiet_stdout q;
        anonymouslibHandle<int, unsigned int, double> A(m, n);
        err = A.inputCSR(nnz, rp, ci, va);
        if (!err) err = A.setX(const_cast<double *>(x));
        A.setSigma(ANONYMOUSLIB_CSR5_SIGMA);
        if (!err) err = A.asCSR5();
        if (!err) err = A.spmv(1.0, y);
        A.destroy();
    }
    _mm_free(rp); _mm_free(ci); _mm_free(va);
    return err;
}

// Timing protocol of CSR5_avx2/main.cpp:41-79: 5 asCSR5/asCSR round trips, one timed asCSR5, one
// checked spmv into y, `warmup` spmv's on re-zeroed y, then `runs` back-to-back spmv's timed with
// gettimeofday.  Outputs mean ms per SpMV and the conversion time.
int ref_avx2_bench(int m, int n, int nnz, const int *row_ptr, const int *col, const double *val,
                   const double *x, double *y, int nthreads, int warmup, int runs,
                   double *ms_per_spmv, double *convert_ms)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    int *rp = (int *)_mm_malloc((size_t)(m + 2) * sizeof(int), 64);
    int *ci = (int *)_mm_malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int), 64);
    double *va = (double *)_mm_malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(double), 64);
    double *yb = (double *)_mm_malloc((size_t)m * sizeof(double), 64);
    memcpy(rp, row_ptr, (size_t)(m + 1) * sizeof(int));
    rp[m + 1] = rp[m];
    memcpy(ci, col, (size_t)nnz * sizeof(int));
    memcpy(va, val, (size_t)nnz * sizeof(double));
    memset(y, 0, (size_t)m * sizeof(double));
    memset(yb, 0, (size_t)m * sizeof(double));

    int err;
    {
        quiet_stdout q;
        anonymouslibHandle<int, unsigned int, double> A(m, n);
        err = A.inputCSR(nnz, rp, ci, va);
        if (!err) err = A.setX(const_cast<double *>(x));
        A.setSigma(ANONYMOUSLIB_CSR5_SIGMA);
        double t0 = now_ms();
        if (!err) err = A.asCSR5();
        *convert_ms = now_ms() - t0;
        if (!err) err = A.spmv(1.0, y);
        for (int i = 0; i < warmup && !err; i++) {
            memset(yb, 0, (size_t)m * sizeof(double));
            err = A.spmv(1.0, yb);
        }
        t0 = now_ms();
        for (int i = 0; i < runs && !err; i++) err = A.spmv(1.0, yb);
        *ms_per_spmv = runs > 0 ? (now_ms() - t0) / runs : 0.0;
        A.destroy();
    }
    _mm_free(rp); _mm_free(ci); _mm_free(va); _mm_free(yb);
    return err;
}

// FP32 has no AVX2 path in the reference (README.md:36).  CPU baseline for the FP32 config is the
// reference's scalar CSR loop (CSR5_cuda/main.cu:343-349) spread over rows with OpenMP ("port").
int ref_csr_omp_bench_f32(int m, const int *row_ptr, const int *col, const float *val,
                          const float *x, float *y, int nthreads, int warmup, int runs,
                          double *ms_per_spmv)
{
    if (nthreads > 0) omp_set_num_threads(nthreads);
    double t0 = 0;
    for (int it = 0; it < warmup + runs; it++) {
        if (it == warmup) t0 = now_ms();
#pragma omp parallel for schedule(static)
        for (int i = 0; i < m; i++) {
            float sum = 0;
            for (int j = row_ptr[i]; j < row_ptr[i + 1]; j++) sum += x[col[j]] * val[j];
            y[i] = sum;
        }
    }
    *ms_per_spmv = runs > 0 ? (now_ms() - t0) / runs : 0.0;
    return 0;
}

}  // extern "C"
