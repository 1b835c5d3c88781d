/*
 * csr5_b200.h -- C ABI of libcsr5_b200.so, the B200-native (sm_100a) CSR5 SpMV.
 *
 * This is the drop-in boundary for the hot path of weifengliu-ssslab/Benchmark_SpMV_using_CSR5:
 * one entry point per public method of the reference's `anonymouslibHandle<int, unsigned int, VT>`
 * (CSR5_cuda/anonymouslib_cuda.h:11-24), with the template parameter VT folded into `value_bytes`
 * (8 = double, 4 = float; VALUE_TYPE in CSR5_cuda/Makefile:4).  include/anonymouslib_cuda.h wraps
 * these symbols back into the reference's class template so CSR5_cuda/main.cu:59-108 compiles
 * unchanged; INTEGRATION.md shows that binding and the ctypes one.
 *
 * Conventions shared with the reference:
 *   - all array arguments are DEVICE pointers, borrowed, int32 indices, 0-based;
 *   - `col` / `val` are permuted IN PLACE by as_csr5() and restored by as_csr() / destroy()
 *     (anonymouslib_cuda.h:79-103, 204-205);
 *   - return value: ANONYMOUSLIB_* codes of detail/common.h:13-18, 0 = success;
 *   - not thread-safe per handle; everything is issued on one stream (default: the legacy
 *     default stream, like the reference); as_csr5()/as_csr() are synchronous, spmv() is
 *     asynchronous.
 *
 * Deliberate differences, each a superset of the reference's behaviour on its own valid inputs
 * (y zeroed, alpha == 1):
 *   - spmv() OVERWRITES y: every row of y is written on every call (empty rows get 0), so y need
 *     not be zeroed and repeated calls are idempotent (the reference accumulates into rows at
 *     tile starts and drifts, CSR5_cuda/main.cu:84-99);
 *   - alpha is honoured (the reference's kernels ignore it, csr5_spmv_cuda.h:22, while its own
 *     scalar check multiplies by it, main.cu:347); CSR5B200_OPT_IGNORE_ALPHA restores the bug;
 *   - CUDA errors are returned (CSR5B200_CUDA_ERROR) instead of exit()ing; sigma outside [4, 32]
 *     is rejected by as_csr5() instead of silently launching nothing.
 */
#ifndef CSR5_B200_H
#define CSR5_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define CSR5B200_API __attribute__((visibility("default")))
#else
#define CSR5B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* detail/common.h:13-18 */
#define CSR5B200_SUCCESS                  0
#define CSR5B200_UNKNOWN_FORMAT          (-1)
#define CSR5B200_UNSUPPORTED_CSR5_OMEGA  (-2)
#define CSR5B200_CSR_TO_CSR5_FAILED      (-3)
#define CSR5B200_UNSUPPORTED_CSR_SPMV    (-4)
#define CSR5B200_UNSUPPORTED_VALUE_TYPE  (-5)
/* not in the reference (it aborts through checkCudaErrors instead) */
#define CSR5B200_CUDA_ERROR              (-100)
#define CSR5B200_INVALID_ARGUMENT        (-101)

/* detail/common.h:20-22 */
#define CSR5B200_FORMAT_CSR   0
#define CSR5B200_FORMAT_CSR5  1

/* detail/cuda/common_cuda.h:11,15 */
#define CSR5B200_OMEGA             32
#define CSR5B200_AUTO_TUNED_SIGMA  (-1)

typedef struct csr5b200_handle_s *csr5b200_handle_t;

/* ---- the reference's public methods ---------------------------------------------------------- */

/* anonymouslibHandle(m, n)                                   anonymouslib_cuda.h:15 */
CSR5B200_API int csr5b200_create(int m, int n, int value_bytes, csr5b200_handle_t *out);
/* int warmup()                                               anonymouslib_cuda.h:16, 55-59 */
CSR5B200_API int csr5b200_warmup(csr5b200_handle_t h);
/* int inputCSR(nnz, row_pointer, column_index, value)        anonymouslib_cuda.h:17, 61-76 */
CSR5B200_API int csr5b200_input_csr(csr5b200_handle_t h, int nnz, int *row_ptr, int *col, void *val);
/* int asCSR()                                                anonymouslib_cuda.h:18, 78-103 */
CSR5B200_API int csr5b200_as_csr(csr5b200_handle_t h);
/* int asCSR5()                                               anonymouslib_cuda.h:19, 105-220 */
CSR5B200_API int csr5b200_as_csr5(csr5b200_handle_t h);
/* int setX(x)                                                anonymouslib_cuda.h:20, 222-260 */
CSR5B200_API int csr5b200_set_x(csr5b200_handle_t h, void *x);
/* int spmv(alpha, y)                                         anonymouslib_cuda.h:21, 262-285 */
CSR5B200_API int csr5b200_spmv(csr5b200_handle_t h, double alpha, void *y);
/* int destroy()   (asCSR + release of the CSR5 arrays)       anonymouslib_cuda.h:22, 287-292 */
CSR5B200_API int csr5b200_destroy(csr5b200_handle_t h);
/* void setSigma(sigma | ANONYMOUSLIB_AUTO_TUNED_SIGMA)       anonymouslib_cuda.h:23, 294-318 */
CSR5B200_API int csr5b200_set_sigma(csr5b200_handle_t h, int sigma);

/* ---- multi-GPU (no reference counterpart: the reference is single-device) -------- */

#define CSR5B200_MAX_SCATTER 8

/* spmv() of a row-range shard that also delivers the result to the other GPUs -- the all-gather of the
 * y segments fused into the SpMV kernels.
 *   y_local : this shard's y segment in LOCAL device memory; carries are accumulated here, and on return of
 *             the enqueued work it holds the final rows like after spmv().
 *   y_dst[k]: address of THIS shard's first row inside destination k's concatenated y (device pointers;
 *             peers' buffers mapped into this process over NVLink, e.g. CUDA IPC / symmetric memory).  The
 *             list normally contains y_local itself.  With dst_is_multicast = 1, n_dst must be 1 and
 *             y_dst[0] is an NVSwitch multicast address (multimem) covering all GPUs including this one.
 * Two exchange schemes (CSR5B200_OPT_EXCHANGE): "fused" -- the SpMV kernels store every finished row to all
 * destinations as tiles complete, so the NVLink traffic overlaps the tile stream; rows completed by carries
 * are re-sent once after the local carry pass (plain stores only -- no atomics cross NVLink), empty rows are
 * cleared everywhere; "push" -- the SpMV runs on local memory and one coalesced pass then copies the segment
 * to every destination with 16-byte stores (for matrices whose row stores are scattered).  Auto picks fused
 * for matrices without empty rows and with short rows.  Everything is ordered on the handle's stream; the
 * caller synchronises the devices afterwards (a cross-GPU barrier) before anyone reads y.
 * 1 <= n_dst <= CSR5B200_MAX_SCATTER.  Same return codes as spmv(). */
CSR5B200_API int csr5b200_spmv_scatter(csr5b200_handle_t h, double alpha, void *y_local, int n_dst,
                                       void *const *y_dst, int dst_is_multicast);

/* Frees the handle object itself (the reference's handle is a stack object). Calls destroy(). */
CSR5B200_API int csr5b200_free(csr5b200_handle_t h);

/* ---- additions that have no reference counterpart -------------------------------------------- */

/* Stream all later work of this handle is issued on (a cudaStream_t; NULL = legacy default). */
CSR5B200_API int csr5b200_set_stream(csr5b200_handle_t h, void *cuda_stream);

#define CSR5B200_OPT_KERNEL        1  /* 0 auto (default; = the faster one on B200: direct-load), 1 direct-load kernel, 2 TMA-staged kernel,
                                         4 TMA-staged kernel with the x gathers prefetched one tile ahead */
#define CSR5B200_OPT_IGNORE_ALPHA  2  /* 1 = reference bug-compat: alpha treated as 1 */
#define CSR5B200_OPT_TMA_STAGES    3  /* smem ring depth of the TMA-staged kernel (0 = default) */
#define CSR5B200_OPT_TMA_WARPS     4  /* consumer warps per CTA of the TMA-staged kernel (0 = default) */
#define CSR5B200_OPT_CTAS_PER_SM   5  /* persistent-grid size factor (0 = default) */
#define CSR5B200_OPT_KERNEL_TIMING 6  /* 1 = bracket the main SpMV kernel of every spmv() with CUDA events
                                         (see csr5b200_get_kernel_times); 0 = off (default) */
#define CSR5B200_OPT_DIRECT_WPB    7  /* tuning: warps per CTA of the direct kernel (2/4/8/16; 0 = default) */
#define CSR5B200_OPT_DIRECT_NCH    8  /* tuning: register chunks per tile (1/2/3; 0 = default rule) */
#define CSR5B200_OPT_HOT_COLUMNS    9  /* hot-column table, decided at as_csr5(): 0 off (default), -1 auto (kept when it
                                         serves >= 25 % of the x references), K > 0 = table capacity in entries.
                                         While it is on, the tagged occurrences in `col` read (bit 31 | slot). */
#define CSR5B200_OPT_HOT_THREADS   10 /* tuning: threads per CTA of the hot-column kernel (0 = default 768) */
#define CSR5B200_OPT_EXCHANGE      11 /* spmv_scatter: 0 auto (default), 1 fused (the SpMV kernels store every finished row to
                                         all destinations), 2 push (one coalesced copy pass after the SpMV) */
CSR5B200_API int csr5b200_set_option(csr5b200_handle_t h, int option, int value);

/* Introspection for tests and harnesses: scalars + device pointers of the CSR5 arrays
 * (anonymouslib_cuda.h:25-52).  Pointers are valid while the handle is in CSR5 format. */
typedef struct csr5b200_info {
    int format;            /* CSR5B200_FORMAT_* */
    int m, n, nnz;
    int value_bytes;
    int sigma;             /* _csr5_sigma */
    int bit_y_offset;      /* _bit_y_offset */
    int bit_scansum_offset;/* _bit_scansum_offset */
    int num_packet;        /* _num_packet */
    int p;                 /* _p: number of tiles, the last one is the tail */
    int num_offsets;       /* _num_offsets */
    int tail_partition_start; /* _tail_partition_start */
    int needs_zero_fill;   /* 1 if some row before the tail is empty (y is memset inside spmv) */
    int kernel_in_use;     /* 1 direct-load, 2 TMA-staged, 3 hot-column (direct-load + x table in shared memory), 4 TMA-staged + x prefetch */
    const uint32_t *partition_pointer;           /* (p + 1) */
    const uint32_t *partition_descriptor;        /* p * 32 * num_packet */
    const int32_t  *partition_descriptor_offset_pointer; /* (p + 1) */
    const int32_t  *partition_descriptor_offset; /* num_offsets */
    const void     *calibrator;                  /* p values */
    int last_cuda_error;   /* cudaError_t of the last failing CUDA call, 0 if none */
    int launches_per_spmv; /* kernels (+ memset nodes) one spmv() enqueues */
    int hot_columns;       /* entries of the hot-column table in use (0 = none) */
    double hot_coverage;   /* fraction of the tiles' x references served by the table */
} csr5b200_info;
CSR5B200_API int csr5b200_get_info(csr5b200_handle_t h, csr5b200_info *out);

/* With CSR5B200_OPT_KERNEL_TIMING on: device durations (ms, CUDA events on the handle's stream) of the
 * main SpMV kernel of the spmv() calls since the last call of this function, oldest first, at most
 * `capacity` (and at most 4096 are retained).  Synchronises the stream.  *count = number written. */
CSR5B200_API int csr5b200_get_kernel_times(csr5b200_handle_t h, float *ms, int capacity, int *count);

/* Host copies of the CSR5 arrays (sizes as in csr5b200_info; NULL = skip that array).  Synchronous.
 * Test/diagnostic aid: lets a harness diff the metadata word for word without a CUDA binding. */
CSR5B200_API int csr5b200_copy_meta_to_host(csr5b200_handle_t h, uint32_t *partition_pointer,
                                            uint32_t *partition_descriptor,
                                            int32_t *partition_descriptor_offset_pointer,
                                            int32_t *partition_descriptor_offset, void *calibrator);

/* Host-buffer entry points: the part of the reference's call_anonymouslib() (CSR5_cuda/main.cu:
 * 17-117) that moves data, for callers that hold host arrays.
 *   spmv_host: H2D copy of x (n values), spmv, D2H copy of y (m values), synchronous.
 *   x_host / y_host may be pageable or pinned; the handle owns the device staging buffers. */
CSR5B200_API int csr5b200_spmv_host(csr5b200_handle_t h, double alpha, const void *x_host, void *y_host);

/* `count` independent SpMVs y_k = alpha * A * x_k on host vectors, software-pipelined: the upload of
 * x_{k+1}, the SpMV of x_k and the download of y_{k-1} run concurrently (two device buffers per
 * direction, separate copy streams, PCIe is full duplex).  x_hosts[k] / y_hosts[k] should be pinned for
 * the copies to be asynchronous.  Synchronous: returns when every y_k is in host memory. */
CSR5B200_API int csr5b200_spmv_host_batch(csr5b200_handle_t h, double alpha, int count, const void *const *x_hosts,
                                          void *const *y_hosts);

/* One-shot equivalent of call_anonymouslib() without the benchmark loop: uploads the CSR arrays
 * and x, converts, runs ONE spmv, downloads y, restores and frees everything. */
CSR5B200_API int csr5b200_call_anonymouslib(int m, int n, int nnz, const int *row_ptr_host, const int *col_host,
                               const void *val_host, const void *x_host, void *y_host,
                               double alpha, int value_bytes, int sigma);

CSR5B200_API const char *csr5b200_version(void);
CSR5B200_API const char *csr5b200_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif /* CSR5_B200_H */
