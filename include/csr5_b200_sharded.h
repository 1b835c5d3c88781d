/*
 * csr5_b200_sharded.h -- C ABI of the row-range sharded CSR5 SpMV over the GPUs of one box, driven from ONE host
 * process (part of libcsr5_b200.so).
 *
 * No reference counterpart: weifengliu-ssslab/Benchmark_SpMV_using_CSR5 is single-device (SURVEY.md s2,
 * "distributed communication backend: none").  This is the C/C++ host API of the multi-GPU form BASELINE.json's
 * north_star asks for ("host code stays C++"; "the matrix shards by tile-partition range across the 8 GPUs of one
 * box with x replicated and y segments concatenated over NVLink"):
 *
 *   - the m x n matrix is cut into contiguous row ranges of balanced nnz: shard g starts at the row that holds
 *     nnz index g * nnz / G, the last such row on ties -- the search of the reference's tile partitioning
 *     (generate_partition_pointer_s1_kernel, CSR5_cuda/detail/cuda/format_cuda.h:21-42) applied to shard
 *     boundaries;
 *   - every shard is an ordinary csr5b200 handle on its device (own CSR5 arrays, own sigma), x is replicated;
 *   - a step is csr5b200_spmv_allgather (csr5_b200.h) on every shard: the SpMV runs in row blocks while the finished
 *     blocks travel to the other GPUs over peer-mapped memory; afterwards EVERY device holds all of y.
 *     y is double-buffered, so steps can be enqueued back to back and the y of one step can be the x of the next
 *     (csr5b200_sharded_iterate) without a host round trip.
 *
 * A device may be listed more than once (several shards on one GPU): that is how the single-GPU test box
 * exercises the whole path.  One worker thread per shard issues that shard's CUDA work.  The multi-process form of
 * the same step (one process per GPU, torch.distributed + symmetric memory) is benchmark_spmv_using_csr5_b200/
 * sharded.py; both call the same csr5b200_spmv_allgather.
 *
 * Return codes as in csr5_b200.h.  Not thread-safe per object.
 */
#ifndef CSR5_B200_SHARDED_H
#define CSR5_B200_SHARDED_H

#include "csr5_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct csr5b200_sharded_s *csr5b200_sharded_t;

/* How the shards agree that a step is complete. */
#define CSR5B200_BARRIER_AUTO    0  /* FLAGS when every shard has its own device, else EVENTS */
#define CSR5B200_BARRIER_FLAGS   1  /* device-side: flag words in peer memory (csr5b200_spmv_allgather's own barriers) */
#define CSR5B200_BARRIER_EVENTS  2  /* host-enqueued: every shard's stream waits for every other shard's event */

/* n_shards <= CSR5B200_MAX_SCATTER; devices[s] = CUDA device ordinal of shard s.  Enables peer access. */
CSR5B200_API int csr5b200_sharded_create(int n_shards, const int *devices, int value_bytes, csr5b200_sharded_t *out);
/* Host CSR of the whole matrix (int32, 0-based; as the reference's loader produces it, main.cu:211-306): split by
 * balanced nnz and upload each row range to its device.  The host arrays are not kept. */
CSR5B200_API int csr5b200_sharded_input_csr_host(csr5b200_sharded_t s, int m, int n, int nnz, const int *row_ptr,
                                                 const int *col, const void *val);
/* Before input_csr_host: with row_cost > 0 the row ranges minimise max over shards of max(nnz, row_cost * rows)
 * instead of balancing nnz alone (default 0).  A shard's SpMV costs its non-zeros; its share of the y exchange, which
 * runs while the SpMV runs, costs its rows -- row_cost non-zeros' worth each (8 on B200: 8 bytes at the ~0.25 TB/s one
 * GPU's multicast stream sustains against 12 bytes per non-zero at half the HBM rate).  Where rows are not the
 * bottleneck this is the nnz rule; a power-law matrix on many GPUs sheds rows from the shard that holds the short and
 * empty ones (R-MAT 25 on 8 GPUs by nnz alone: 14.6 M of 33.5 M rows in the last shard; 0.70 -> 0.58 ms per step). */
CSR5B200_API int csr5b200_sharded_set_partition(csr5b200_sharded_t s, double row_cost);
/* sigma for every shard (CSR5B200_AUTO_TUNED_SIGMA: each shard applies the rule to its own nnz / m). */
CSR5B200_API int csr5b200_sharded_set_sigma(csr5b200_sharded_t s, int sigma);
/* csr5b200_set_option on every shard's handle. */
CSR5B200_API int csr5b200_sharded_set_option(csr5b200_sharded_t s, int option, int value);
/* transport: CSR5B200_TRANSPORT_* (multicast needs the multi-process binding); chunks, push_ctas: 0 = default;
 * barrier: CSR5B200_BARRIER_*. */
CSR5B200_API int csr5b200_sharded_set_exchange(csr5b200_sharded_t s, int transport, int chunks, int push_ctas,
                                               int barrier, int timeout_ms);
/* Replicate x (n values) on every device. */
CSR5B200_API int csr5b200_sharded_set_x_host(csr5b200_sharded_t s, const void *x_host);
CSR5B200_API int csr5b200_sharded_as_csr5(csr5b200_sharded_t s);
/* One step, asynchronous: y = alpha * A x + beta * y on every device. */
CSR5B200_API int csr5b200_sharded_spmv(csr5b200_sharded_t s, double alpha, double beta);
/* `steps` steps of x <- alpha * A x on the devices (square matrix); afterwards x and the current y are the last
 * iterate.  Asynchronous. */
CSR5B200_API int csr5b200_sharded_iterate(csr5b200_sharded_t s, int steps, double alpha);
/* Waits for all shards; CSR5B200_EXCHANGE_TIMEOUT if a device-side barrier gave up. */
CSR5B200_API int csr5b200_sharded_synchronize(csr5b200_sharded_t s);
/* Device pointer of the concatenated y (m values) of the last step on shard `shard`'s device. */
CSR5B200_API int csr5b200_sharded_get_y(csr5b200_sharded_t s, int shard, void **y_dev);
/* Synchronises, then copies that y to host memory. */
CSR5B200_API int csr5b200_sharded_copy_y_to_host(csr5b200_sharded_t s, int shard, void *y_host);
/* bounds[0 .. n_shards]: first row of every shard, bounds[n_shards] = m. */
CSR5B200_API int csr5b200_sharded_get_bounds(csr5b200_sharded_t s, long long *bounds);
/* The ordinary handle of one shard (introspection: csr5b200_get_info, csr5b200_copy_meta_to_host). */
CSR5B200_API int csr5b200_sharded_get_handle(csr5b200_sharded_t s, int shard, csr5b200_handle_t *h);
/* Restores nothing on the host (the device copies are owned here); frees everything. */
CSR5B200_API int csr5b200_sharded_destroy(csr5b200_sharded_t s);

#ifdef __cplusplus
}
#endif
#endif /* CSR5_B200_SHARDED_H */
