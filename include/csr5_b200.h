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
/* y = alpha * A * x + beta * y.  The reference's spmv() carries the stub of this form -- a commented-out `beta`
 * argument in the csr5_spmv() call at anonymouslib_cuda.h:281 -- and never implements it.  beta = 0 is exactly spmv() (y is not read). */
CSR5B200_API int csr5b200_spmv_axpby(csr5b200_handle_t h, double alpha, double beta, void *y);
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
 *             list contains y_local itself (required by the fused scheme).  With dst_is_multicast = 1, n_dst must be 1 and
 *             y_dst[0] is an NVSwitch multicast address (multimem) covering all GPUs including this one.
 * Two exchange schemes (CSR5B200_OPT_EXCHANGE): "fused" -- the SpMV kernels store every finished row to all
 * destinations as tiles complete, so the NVLink traffic overlaps the tile stream; rows completed by carries
 * are re-sent once after the local carry pass (plain stores only -- no atomics cross NVLink), empty rows are
 * cleared everywhere; "push" -- the SpMV runs on local memory and one coalesced pass then copies the segment
 * to every destination with 16-byte stores (for matrices whose row stores are scattered).  Auto picks fused
 * for matrices without empty rows and with short rows.  Everything is ordered on the handle's stream; the
 * caller synchronises the devices afterwards (a cross-GPU barrier) before anyone reads y -- and BEFORE the call
 * too if a peer may still be reading the y of the previous step (the stores land in the peers' buffers as soon as
 * the kernel runs).  In the fused scheme y_dst[] MUST contain y_local (carries are completed in y_local and
 * re-sent from there), else CSR5B200_INVALID_ARGUMENT.  Superseded by csr5b200_spmv_allgather below, which
 * overlaps the exchange with the SpMV and brings its own barriers.
 * 1 <= n_dst <= CSR5B200_MAX_SCATTER.  Same return codes as spmv(). */
CSR5B200_API int csr5b200_spmv_scatter(csr5b200_handle_t h, double alpha, void *y_local, int n_dst,
                                       void *const *y_dst, int dst_is_multicast);

/* ---- overlapped all-gather: the step of a row-range sharded SpMV, one call per shard per step -----------------
 *
 * The shard's CSR5 tiles are cut into `chunks` row blocks.  The blocks stream through the SpMV kernel back to back
 * (two alternating streams, so a block's last CTAs overlap the next block's first); as soon as a block's carries
 * are in, its finished rows travel to every other GPU while the later blocks are still being computed:
 *   COPY_ENGINE     one peer copy per destination on its own stream (DMA engines; no SM is involved),
 *   SM_PUSH         a small grid (push_ctas CTAs) of 16-byte coalesced loads + one store per peer,
 *   SM_MULTICAST    the same grid storing once to the NVSwitch multicast address (multimem.st),
 *   IN_KERNEL       no row blocks: the SpMV kernel itself stores each finished row to every destination
 *                   (the first-generation scheme of csr5b200_spmv_scatter; beta must be 0),
 *   NONE            nothing is sent (y_full[rank] only receives this shard's rows).
 * Two device-side barriers on flag words in peer-mapped memory bracket the remote writes: the ENTRY barrier
 * (overlapped with the first row block) holds them back until every rank has finished the work it had enqueued
 * before this call -- e.g. reading the previous y -- and the EXIT barrier ends the step: after it, on this
 * handle's stream, y_full[rank] holds all rows of all shards.  Callers that alternate between two y buffers can
 * drop the entry barrier (entry_barrier = 0); callers that synchronise the ranks themselves pass flags = NULLs.
 * Every rank must make the same sequence of calls.  Barriers give up after timeout_ms (default 20 s) and
 * csr5b200_exchange_status() then reports it -- a lost peer never hangs the GPU. */
#define CSR5B200_TRANSPORT_AUTO         0
#define CSR5B200_TRANSPORT_COPY_ENGINE  1
#define CSR5B200_TRANSPORT_SM_PUSH      2
#define CSR5B200_TRANSPORT_SM_MULTICAST 3
#define CSR5B200_TRANSPORT_IN_KERNEL    4
#define CSR5B200_TRANSPORT_NONE         5

typedef struct csr5b200_exchange {
    int rank, world;                        /* 1 <= world <= CSR5B200_MAX_SCATTER */
    void *y_full[CSR5B200_MAX_SCATTER];     /* base of rank k's concatenated y as mapped in THIS process; [rank] is local */
    void *y_multicast;                      /* NVSwitch multicast address of the same buffer, or NULL */
    uint32_t *flags[CSR5B200_MAX_SCATTER];  /* rank k's barrier words (>= 2 * world, zeroed once) as mapped here; NULL = no barriers */
    long long row_begin;                    /* first row of this shard inside the concatenated y */
    int chunks;                             /* row blocks (0 = default: 12, or 8 for shards whose blocks differ widely in rows; at most 64) */
    int transport;                          /* CSR5B200_TRANSPORT_* */
    int entry_barrier;                      /* 1 = hold remote writes until all ranks reached this call */
    int push_ctas;                          /* SM transports: CTAs of the push grid (0 = default 48) */
    int timeout_ms;                         /* barrier time-out (0 = default 20000) */
    int push_threads;                       /* SM transports: threads per CTA of the push grid (0 = default 256, at most 1024) */
} csr5b200_exchange;

/* y_full[rank][row_begin + i] = alpha * (A_shard x)_i + beta * (old value), i < m, delivered to every rank as
 * described above.  Same return codes as spmv(). */
CSR5B200_API int csr5b200_spmv_allgather(csr5b200_handle_t h, double alpha, double beta, const csr5b200_exchange *ex);
/* With CSR5B200_OPT_EXCHANGE_TRACE on: the timeline of the last step on this shard, in ms since the step began, three
 * values per row block in execution order -- SpMV tiles done, carry pass done, rows shipped (-1: nothing shipped by an
 * SM transport) -- and finally the end of the step (after the exit barrier).  Synchronises the stream. */
CSR5B200_API int csr5b200_exchange_trace(csr5b200_handle_t h, float *ms, int capacity, int *count);
/* Synchronises the handle's stream; returns CSR5B200_EXCHANGE_TIMEOUT if a barrier gave up since the last call. */
#define CSR5B200_EXCHANGE_TIMEOUT (-102)
CSR5B200_API int csr5b200_exchange_status(csr5b200_handle_t h);

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
#define CSR5B200_OPT_SIGMA_RULE    12 /* what CSR5B200_AUTO_TUNED_SIGMA means at the next set_sigma(): 0 (default) the reference's
                                         table r/s/t/u = 4/32/256/6 (anonymouslib_cuda.h:297-313; keeps the CSR5 arrays word for
                                         word those of the reference), 1 the rule measured on B200 (profiles/r02_sigma_rule.md); the environment variable
                                         CSR5B200_SIGMA_RULE=b200 sets it for every handle of an unmodified caller */
#define CSR5B200_OPT_DETERMINISTIC 13 /* 1 = spmv() / spmv_axpby() add the carries of a row in tile order, one thread per row, instead
                                         of with floating-point atomics (replaces csr5_spmv_cuda.h:313-382, which mixes both): the
                                         result is bit-identical from run to run also for rows that span several tiles; 0 = default */
#define CSR5B200_OPT_EXCHANGE_TRACE 14 /* 1 = csr5b200_spmv_allgather records timing events per row block (csr5b200_exchange_trace) */
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
    float convert_phase_ms[8]; /* device time of the phases of the last as_csr5(): [0] tile_ptr, [1] tile_desc,
                                  [2] scan, [3] desc_offset, [4] transpose (in-place, col + val) */
    double convert_host_ms;    /* host wall time of the last as_csr5() */
    double convert_alloc_ms;   /* of which buffer (re)allocation; 0 when the handle's buffer pool already fits */
    int exchange_transport;    /* csr5b200_spmv_allgather: transport and row blocks of the last step */
    int exchange_chunks;
    int has_carries;           /* 0: every tile starts on a row boundary, spmv() is ONE launch (no carry pass) */
} csr5b200_info;
CSR5B200_API int csr5b200_get_info(csr5b200_handle_t h, csr5b200_info *out);

/* With CSR5B200_OPT_KERNEL_TIMING on: device durations (ms, CUDA events on the handle's stream) of the
 * main SpMV kernel of the spmv() calls since the last call of this function, oldest first, at most
 * `capacity` (and at most 4096 are retained).  Synchronises the stream.  *count = number written. */
CSR5B200_API int csr5b200_get_kernel_times(csr5b200_handle_t h, float *ms, int capacity, int *count);

/* COO -> CSR on the device with the semantics of the reference's loader (CSR5_cuda/main.cu:211-306, the same
 * loops in every backend's main): 0-based (row, col[, val]) triples in file order; with symmetric != 0 every
 * off-diagonal (i, j) also yields (j, i) right after it (main.cu:239-246, 271-289; the reference does this for
 * symmetric and hermitian files, not for skew-symmetric ones); rows are filled in emission order -- a STABLE sort by
 * row: columns are not sorted inside a row, duplicates are kept.  vals == NULL: pattern file, every value is 1.
 * All pointers are device pointers except nnz_out (host).  row_ptr has m + 1 entries; col_out / val_out have
 * `capacity` entries.  *nnz_out = entries of the CSR (nnz + mirrored ones); call with col_out == NULL to only
 * count (row_ptr is scratch then).  CSR5B200_INVALID_ARGUMENT: an index is out of range, or capacity is too small.
 * Synchronous on `cuda_stream`. */
CSR5B200_API int csr5b200_coo_to_csr(int m, int n, int nnz, const int *rows, const int *cols, const void *vals,
                                     int value_bytes, int symmetric, int *row_ptr, int *col_out, void *val_out,
                                     int capacity, int *nnz_out, void *cuda_stream);

/* Microbenchmarks on the matrix held by the handle (CSR5 format, no hot-column table): the SpMV's launch shape and
 * col stream with the kernel stripped down to one component, to MEASURE the floors quoted in DESIGN.md
 * (csrc/csr5_probe.cu).  *ms_avg = mean device time of `repeats` launches.  Overwrites the calibrator array. */
#define CSR5B200_PROBE_STREAM       1  /* val + col streamed, no x: the HBM floor of the matrix stream */
#define CSR5B200_PROBE_GATHER_NC    2  /* col streamed + x[col] through ld.global.nc: the gather floor as the kernel issues it */
#define CSR5B200_PROBE_GATHER_CG    3  /* ... through ld.global.cg */
#define CSR5B200_PROBE_GATHER_CA    4  /* ... through ld.global.ca */
#define CSR5B200_PROBE_GATHER_ONLY  5  /* no col stream: x[hash]: the pure divergent-gather rate out of L2 */
#define CSR5B200_PROBE_FMA_NOSEG    6  /* val + col + x + FMA, no descriptors, no segmented sum, one store per tile */
CSR5B200_API int csr5b200_probe(csr5b200_handle_t h, int kind, int repeats, float *ms_avg);

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
