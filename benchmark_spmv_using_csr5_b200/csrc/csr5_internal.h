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
    int has_carries = 1;    // 0: every tile (and the tail) starts on a row boundary -- the carry pass has nothing to add

    const int *row_ptr = nullptr;   // borrowed
    int *col = nullptr;             // borrowed, permuted in place while in CSR5 format
    void *val = nullptr;            // borrowed, permuted in place while in CSR5 format
    const void *x = nullptr;        // borrowed

    uint32_t *tile_ptr = nullptr;   // owned, p + 1
    uint32_t *desc = nullptr;       // owned, p * 32 * num_packet
    int *desc_off_ptr = nullptr;    // owned, p + 1
    int *desc_off = nullptr;        // owned, num_offsets
    void *calibrator = nullptr;     // owned, p values
    int *dev_flags = nullptr;       // owned, small scratch: [0] any dirty tile before the tail, [1] any tile that continues a row

    // Hot-column table (no reference counterpart).  While hot_k > 0 the column indices
    // of the hot_k most referenced columns are stored in the CSR5 tiles as (bit 31 | slot); asCSR()
    // restores them.  The SpMV stages x[hot_col[*]] in shared memory.
    int hot_k = 0;
    int *hot_col = nullptr;         // owned, hot_k entries: slot -> column
    void *hot_x = nullptr;          // owned, hot_k values (16-byte padded): x[hot_col[slot]], refreshed per SpMV
    double hot_coverage = 0.0;      // fraction of the tiles' column references that hit the table
};

struct SpmvTuning {
    int kernel = 0;        // 0 auto, 1 direct-load, 2 TMA-staged
    int tma_stages = 0;
    int tma_warps = 0;
    int ctas_per_sm = 0;
    int num_sms = 148;
    int direct_wpb = 0;    // tuning: warps per CTA of the direct kernel (0 = default 4)
    int direct_nch = 0;    // tuning: register chunks per tile (0 = default rule)
    int hot_columns = 0;   // 0 off (default), -1 auto (kept if it serves >= 25 % of the references), > 0 capacity
    int hot_threads = 0;   // tuning: threads per CTA of the hot-column kernel (0 = default)
    int exchange = 0;      // sharded mode: 0 auto, 1 fused, 2 push
    int deterministic = 0; // carries of a row are summed in tile order by one thread (no atomics): run-to-run identical bits
    cudaEvent_t ev_begin = nullptr;  // optional: recorded right before / after the main SpMV kernel
    cudaEvent_t ev_end = nullptr;
};

// ---- format conversion (csr5_format.cu) ------------------------------------------------------
// tile_ptr / descriptor / segment counts; leaves the exclusive scan and offset table to the next
// two calls.  All asynchronous on `stream`.
cudaError_t launch_tile_ptr(const Plan &pl, cudaStream_t stream);
cudaError_t launch_tile_desc(const Plan &pl, cudaStream_t stream);
cudaError_t launch_scan_offsets(const Plan &pl, void *scratch, size_t scratch_bytes, cudaStream_t stream);
size_t scan_scratch_bytes(int p);   // for an exclusive scan of p + 1 ints
cudaError_t launch_exclusive_scan(int *data, int n, void *scratch, size_t scratch_bytes, cudaStream_t stream);
// hot-column table construction / removal
cudaError_t launch_hot_count(const int *col, long long limit, int *cnt, int num_sms, cudaStream_t stream);
cudaError_t launch_hot_count_ge(const int *cnt, int n, int threshold, unsigned long long *out2, int num_sms,
                                cudaStream_t stream);
int hot_hist_bins();
cudaError_t launch_hot_hist(const int *cnt, int n, unsigned int *cols, unsigned long long *refs, int num_sms,
                            cudaStream_t stream);
cudaError_t launch_hot_flags(const int *cnt, int n, int threshold, int *slot, cudaStream_t stream);
cudaError_t launch_hot_assign(const int *cnt, int n, int threshold, const int *slot, int *hot_col, int *col,
                              long long limit, int num_sms, cudaStream_t stream);
cudaError_t launch_hot_restore(int *col, long long limit, const int *hot_col, int num_sms, cudaStream_t stream);
cudaError_t launch_desc_offset(const Plan &pl, cudaStream_t stream);
// in-place omega x sigma tile transpose of col and val; r2c = CSR -> CSR5.
cudaError_t launch_transpose(const Plan &pl, bool r2c, cudaStream_t stream);
cudaError_t launch_warmup(cudaStream_t stream);

// ---- SpMV (csr5_spmv_f64.cu / csr5_spmv_f32.cu via csr5_spmv.cuh) -----------------------------
// Sharded (multi-GPU) mode of one spmv() (csr5b200_spmv_scatter): where the rows go besides local HBM.
struct ShardCtx {
    int n_dst = 0;
    void *const *y_dst = nullptr;   // this rank's segment inside each GPU's concatenated y, or one multicast address
    int multicast = 0;
    int exchange = 0;               // 0 auto, 1 fused (SpMV kernels store to every destination), 2 push (copy pass after)
};

// Row blocks of the overlapped exchange: block c = tiles [tile_begin[c], tile_begin[c + 1]); skip_row[c] = the row
// that crosses INTO block c from an earlier block (its carries are held back for the boundary pass), or -1.
constexpr int MAX_CHUNKS = 64;
struct ChunkTable {
    int n = 0;
    int tile_begin[MAX_CHUNKS + 1] = {};
    int skip_row[MAX_CHUNKS] = {};
};

// Which parts of an SpMV one launch group enqueues.  A whole spmv() = everything on over all tiles; the
// overlapped multi-GPU exchange (csr5_exchange.cu) cuts the tiles into row blocks.
struct SpmvCall {
    int tile_begin = 0;     // CSR5 tiles [tile_begin, tile_end) of [0, p - 1); tile_end < 0 = up to p - 1
    int tile_end = -1;
    bool tail = true;       // also the rows of the tail tile (p - 1)
    bool prologue = true;   // clear (beta = 0) / scale (beta != 0) the rows no tile stores; refresh the hot-column table
    bool tiles = true;      // the main kernel over the tiles (+ tail)
    bool calibrate = true;  // the carry pass over the same tiles (+ the tail tile's carry)
    int skip_row = -1;      // carry pass: leave out the carries into this row (a row that began in an earlier row block)
    const ChunkTable *boundary = nullptr;   // != nullptr: ONLY the boundary pass -- the held-back carries of all blocks
};

// Enqueues the selected parts of y = alpha * A * x + beta * y; in the legacy sharded mode (sh != nullptr) the rows
// are also delivered to every destination (fused into the kernels, or by a copy pass).  y is the result vector in
// local memory.  Returns the kernel variant used in *used; *launches is incremented per enqueued node.
cudaError_t launch_spmv_part_f64(const Plan &pl, const SpmvTuning &tn, double alpha, double beta, double *y,
                                 const ShardCtx *sh, const SpmvCall &call, cudaStream_t stream, int *used, int *launches);
cudaError_t launch_spmv_part_f32(const Plan &pl, const SpmvTuning &tn, float alpha, float beta, float *y,
                                 const ShardCtx *sh, const SpmvCall &call, cudaStream_t stream, int *used, int *launches);
// SM transport of the overlapped exchange: coalesced copy of `rows` rows from y_local to dst[0..n_dst) (peer
// addresses, or one NVSwitch multicast address) with `grid` CTAs.
cudaError_t launch_push_rows(int value_bytes, const void *y_local, void *const *dst, int n_dst, int multicast,
                             long long rows, int grid, int threads, cudaStream_t stream);

// y_local[rows[i]] -> the same element of every destination, i < n <= MAX_CHUNKS (the rows that cross row-block
// boundaries, final only after the boundary pass).
cudaError_t launch_push_row_list(int value_bytes, const void *y_local, void *const *dst, int n_dst, int multicast,
                                 const int *rows, int n, cudaStream_t stream);

// ---- cross-GPU barrier on flag words in peer-mapped memory (csr5_exchange.cu) ------------------------------
// flags[k] = base of rank k's flag array as mapped here (>= 2 * world words each, zero-initialised).  Rank r
// bumps its own epoch counter (device word), writes it to word [slot * world + r] of every rank, and waits until
// every word [slot * world + k] of its OWN array has reached it.  Bounded: gives up after timeout_ms and raises
// *status (device word) instead of hanging the GPU.
cudaError_t launch_flag_barrier(uint32_t *const *flags, int rank, int world, int slot, uint32_t *epoch,
                                int *status, int timeout_ms, cudaStream_t stream);

}  // namespace csr5

#endif
