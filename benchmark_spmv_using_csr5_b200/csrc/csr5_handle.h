// csr5_handle.h -- the object behind csr5b200_handle_t, shared by the translation units that implement the
// C ABI (csr5_capi.cu: state machine; csr5_exchange.cu: overlapped multi-GPU exchange).  Not part of the ABI.
#ifndef CSR5_HANDLE_H
#define CSR5_HANDLE_H

#include <vector>

#include "csr5_internal.h"

namespace csr5 {

// Streams / events / cached row-block boundaries of csr5b200_spmv_allgather (created at its first call).
struct ExchangeState {
    bool ready = false;
    bool warmed = false;       // every kernel of the step has been launched once (module loading is done)
    cudaStream_t work1 = nullptr;                       // odd row blocks (even ones run on the handle's stream)
    cudaStream_t side = nullptr;                        // carry passes, entry barrier
    cudaStream_t ship = nullptr;                        // SM transports: the push kernels
    cudaStream_t ce[CSR5B200_MAX_SCATTER] = {};         // copy-engine transport: one stream per destination
    cudaEvent_t ev_begin = nullptr, ev_side_done = nullptr, ev_work1_done = nullptr;
    cudaEvent_t ev_chunk[MAX_CHUNKS] = {}, ev_cal[MAX_CHUNKS] = {};
    cudaEvent_t ev_ce_done[CSR5B200_MAX_SCATTER] = {};
    int chunks = 0;                                     // row blocks the cached boundaries are for
    int auto_chunks = 0;                                // what `chunks = 0` resolved to for this matrix (0 = not yet)
    bool reordered = false;                             // the blocks run ship-heavy first (very different row counts)
    std::vector<int> chunk_tile, chunk_row;             // chunks + 1 boundaries: tiles, rows
    std::vector<int> chunk_carried;                     // row chunk_row[c] began before block c (boundary row)
    std::vector<int> chunk_order;                       // execution order of the blocks
    ChunkTable table;                                   // the same for the boundary pass
    int boundary_rows[MAX_CHUNKS] = {};
    int n_boundary_rows = 0;
    cudaEvent_t ev_ship_done = nullptr;
    uint32_t *epoch = nullptr;                          // device: 2 words, barrier epochs of the entry / exit slot
    int *status = nullptr;                              // device: != 0 after a barrier timed out
    int last_transport = 0, last_chunks = 0;
    // CSR5B200_OPT_EXCHANGE_TRACE: timing events of the last step (csr5b200_exchange_trace)
    bool trace = false;
    cudaEvent_t tv_begin = nullptr, tv_end = nullptr;
    cudaEvent_t tv_tiles[MAX_CHUNKS] = {}, tv_cal[MAX_CHUNKS] = {}, tv_ship[MAX_CHUNKS] = {};
    int traced_chunks = 0;
    bool traced_ship[MAX_CHUNKS] = {};
};

}  // namespace csr5

struct csr5b200_handle_s {
    csr5::Plan pl;
    csr5::SpmvTuning tune;
    int format = -1;          // the reference leaves _format unset until inputCSR
    cudaStream_t stream = 0;  // legacy default stream, like the reference
    int ignore_alpha = 0;
    int last_cuda_error = 0;
    int kernel_in_use = 0;
    int launches_per_spmv = 0;
    int sigma_rule = 0;       // CSR5B200_OPT_SIGMA_RULE
    void *x_stage = nullptr;  // device staging for spmv_host
    void *y_stage = nullptr;
    // spmv_host_batch pipeline: double-buffered staging, copy streams, events
    bool batch_ready = false;
    void *xb[2] = {nullptr, nullptr}, *yb[2] = {nullptr, nullptr};
    cudaStream_t s_in = nullptr, s_out = nullptr;
    // sharded mode (spmv_scatter)
    csr5::ShardCtx shard;
    cudaEvent_t e_in[2] = {nullptr, nullptr}, e_comp[2] = {nullptr, nullptr}, e_out[2] = {nullptr, nullptr};
    int kernel_timing = 0;
    std::vector<cudaEvent_t> ev;  // begin/end pairs of the timed main kernels
    size_t ev_used = 0;           // events recorded since the last get_kernel_times()
    csr5::ExchangeState ex;       // overlapped all-gather exchange (csr5_exchange.cu)
    // Buffers of the CSR5 arrays are kept across asCSR() / asCSR5() cycles (the reference frees and reallocates them,
    // anonymouslib_cuda.h:93-97, 142-151: 6 device-synchronising cudaMalloc + cudaFree per conversion); destroy() frees.
    enum { POOL_TILE_PTR, POOL_DESC, POOL_DESC_OFF_PTR, POOL_CALIBRATOR, POOL_FLAGS, POOL_SCAN, POOL_DESC_OFF, POOL_SLOTS };
    void *pool[POOL_SLOTS] = {};
    size_t pool_cap[POOL_SLOTS] = {};
    cudaEvent_t ev_conv[8] = {};  // phase boundaries of as_csr5()
    float convert_ms[8] = {};     // device time of the conversion phases of the last as_csr5() (csr5b200_info)
    double convert_host_ms = 0.0; // host wall time of the last as_csr5()
    double convert_alloc_ms = 0.0;// of which: buffer (re)allocation
};

namespace csr5 {
int handle_cuda_fail(csr5b200_handle_t h, cudaError_t e);
void release_exchange(csr5b200_handle_t h);
}  // namespace csr5

#endif
