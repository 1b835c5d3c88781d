// csr5_exchange.cu -- the step of a row-range sharded SpMV: csr5b200_spmv_allgather (include/csr5_b200.h).
//
// No reference counterpart (the reference is single-device).  What BASELINE.json's north_star
// asks for -- "y segments concatenated over NVLink" -- is an all-gather of (N-1)/N of y INTO every GPU per step,
// as large as the SpMV's own HBM stream once N >= 4, so it must run WHILE the SpMV runs:
//
//   handle stream S   [prologue] [tiles blk 0]      [tiles blk 2]      ...                      [wait all] [exit barrier]
//   work1             .          .     [tiles blk 1]      [tiles blk 3] ...
//   side              [entry barrier]  [carry 0][ship 0] [carry 1][ship 1] ...
//   ce[k] (k != rank) .                        [copy 0 -> k]      [copy 1 -> k] ...     (COPY_ENGINE transport only)
//
// Row blocks are independent of each other and of their order: a row is stored once, by the tile in which it starts;
// a block's carry pass adds the carries of its own tiles EXCEPT those into the one row that crossed in from an earlier
// block.  After block c and its carry pass every row that lies inside the block is final and leaves; the (at most
// one per block) crossing rows get their held-back carries in one boundary pass after the last block and leave
// last.  Consecutive blocks go to two alternating streams and overlap at their edges -- no partial-wave bubble per
// block -- and the blocks run ship-heavy first (most rows per tile: a power-law shard keeps its millions of short
// and empty rows at the end), the order that minimises the makespan of the two-stage compute -> ship pipeline.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>

#include "csr5_handle.h"

namespace csr5 {

namespace {

struct FlagPtrs {
    uint32_t *p[CSR5B200_MAX_SCATTER];
};

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// One CTA of 32 threads; thread k talks to rank k.
__global__ void flag_barrier_kernel(const FlagPtrs fp, const int rank, const int world, const int slot,
                                    uint32_t *epoch, int *status, const unsigned long long timeout_ns)
{
    if (world <= 0) return;   // module warm-up launch
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = epoch[slot] + 1;
        epoch[slot] = s_epoch;
    }
    __syncthreads();
    const uint32_t e = s_epoch;
    const int k = threadIdx.x;
    if (k >= world) return;
    // Everything this GPU enqueued before the barrier (kernels, peer copies) has completed; publish it.
    __threadfence_system();
    uint32_t *theirs = fp.p[k] + slot * world + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(e) : "memory");
    const uint32_t *mine = fp.p[rank] + slot * world + k;
    const unsigned long long t0 = global_ns();
    unsigned backoff = 32;
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int)(v - e) >= 0) break;
        if (global_ns() - t0 > timeout_ns) { atomicExch(status, 1); break; }
        __nanosleep(backoff);
        if (backoff < 1024) backoff <<= 1;
    }
}

}  // namespace

cudaError_t launch_flag_barrier(uint32_t *const *flags, int rank, int world, int slot, uint32_t *epoch, int *status,
                                int timeout_ms, cudaStream_t stream)
{
    FlagPtrs fp;
    for (int k = 0; k < CSR5B200_MAX_SCATTER; k++) fp.p[k] = (flags && k < world) ? flags[k] : nullptr;
    flag_barrier_kernel<<<1, 32, 0, stream>>>(fp, rank, world, slot, epoch, status,
                                              (unsigned long long)(timeout_ms > 0 ? timeout_ms : 20000) * 1000000ull);
    return cudaGetLastError();
}

cudaError_t launch_push_rows_f64(const void *, void *const *, int, int, long long, int, int, cudaStream_t);
cudaError_t launch_push_rows_f32(const void *, void *const *, int, int, long long, int, int, cudaStream_t);

cudaError_t launch_push_row_list_f64(const void *, void *const *, int, int, const int *, int, cudaStream_t);
cudaError_t launch_push_row_list_f32(const void *, void *const *, int, int, const int *, int, cudaStream_t);

cudaError_t launch_push_row_list(int value_bytes, const void *y_local, void *const *dst, int n_dst, int multicast,
                                 const int *rows, int n, cudaStream_t stream)
{
    return value_bytes == 8 ? launch_push_row_list_f64(y_local, dst, n_dst, multicast, rows, n, stream)
                            : launch_push_row_list_f32(y_local, dst, n_dst, multicast, rows, n, stream);
}

cudaError_t launch_push_rows(int value_bytes, const void *y_local, void *const *dst, int n_dst, int multicast,
                             long long rows, int grid, int threads, cudaStream_t stream)
{
    return value_bytes == 8 ? launch_push_rows_f64(y_local, dst, n_dst, multicast, rows, grid, threads, stream)
                            : launch_push_rows_f32(y_local, dst, n_dst, multicast, rows, grid, threads, stream);
}

void release_exchange(csr5b200_handle_t h)
{
    ExchangeState &x = h->ex;
    auto ds = [](cudaStream_t &s) { if (s) cudaStreamDestroy(s); s = nullptr; };
    auto de = [](cudaEvent_t &e) { if (e) cudaEventDestroy(e); e = nullptr; };
    ds(x.work1);
    ds(x.side);
    ds(x.ship);
    for (auto &s : x.ce) ds(s);
    de(x.ev_begin);
    de(x.ev_side_done);
    de(x.ev_work1_done);
    de(x.ev_ship_done);
    for (auto &e : x.ev_chunk) de(e);
    for (auto &e : x.ev_cal) de(e);
    for (auto &e : x.ev_ce_done) de(e);
    de(x.tv_begin);
    de(x.tv_end);
    for (auto &e : x.tv_tiles) de(e);
    for (auto &e : x.tv_cal) de(e);
    for (auto &e : x.tv_ship) de(e);
    cudaFree(x.epoch);
    cudaFree(x.status);
    x = ExchangeState();
}

namespace {

#define CUX(call)                                                  \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return handle_cuda_fail(h, e__);   \
    } while (0)

int ensure_state(csr5b200_handle_t h)
{
    ExchangeState &x = h->ex;
    if (x.ready) return CSR5B200_SUCCESS;
    int lo = 0, hi = 0;
    CUX(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = greatest priority (numerically lowest)
    CUX(cudaStreamCreateWithPriority(&x.work1, cudaStreamNonBlocking, lo));
    // the shipping stream outranks the SpMV so that its small kernels get SM slots as soon as CTAs retire
    CUX(cudaStreamCreateWithPriority(&x.side, cudaStreamNonBlocking, hi));
    CUX(cudaStreamCreateWithPriority(&x.ship, cudaStreamNonBlocking, hi));
    for (auto &s : x.ce) CUX(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, hi));
    const unsigned fl = cudaEventDisableTiming;
    CUX(cudaEventCreateWithFlags(&x.ev_begin, fl));
    CUX(cudaEventCreateWithFlags(&x.ev_side_done, fl));
    CUX(cudaEventCreateWithFlags(&x.ev_work1_done, fl));
    CUX(cudaEventCreateWithFlags(&x.ev_ship_done, fl));
    for (auto &e : x.ev_chunk) CUX(cudaEventCreateWithFlags(&e, fl));
    for (auto &e : x.ev_cal) CUX(cudaEventCreateWithFlags(&e, fl));
    for (auto &e : x.ev_ce_done) CUX(cudaEventCreateWithFlags(&e, fl));
    CUX(cudaMalloc(&x.epoch, 2 * sizeof(uint32_t)));
    CUX(cudaMalloc(&x.status, sizeof(int)));
    CUX(cudaMemset(x.epoch, 0, 2 * sizeof(uint32_t)));
    CUX(cudaMemset(x.status, 0, sizeof(int)));
    x.ready = true;
    return CSR5B200_SUCCESS;
}

// Row-block plan (cached): boundaries in tiles, the rows they fall into, which of those rows cross in from the block
// before, and the order the blocks run in.  Two small blocking read-backs.
int ensure_chunks(csr5b200_handle_t h, int chunks)
{
    ExchangeState &x = h->ex;
    const Plan &pl = h->pl;
    const int ntiles = pl.p > 0 ? pl.p - 1 : 0;
    if (chunks > MAX_CHUNKS) chunks = MAX_CHUNKS;
    if (chunks > ntiles) chunks = ntiles;
    if (chunks < 1) chunks = 1;
    if (x.chunks == chunks && (int)x.chunk_tile.size() == chunks + 1) return CSR5B200_SUCCESS;
    x.chunk_tile.assign(chunks + 1, 0);
    x.chunk_row.assign(chunks + 1, 0);
    x.chunk_carried.assign(chunks, 0);
    for (int c = 0; c <= chunks; c++) x.chunk_tile[c] = (int)((long long)ntiles * c / chunks);
    x.chunk_row[chunks] = pl.m;
    std::vector<uint32_t> tp(chunks + 1, 0);
    std::vector<int> rp(chunks + 1, 0);
    for (int c = 0; c < chunks; c++)
        CUX(cudaMemcpyAsync(&tp[c], pl.tile_ptr + x.chunk_tile[c], sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CUX(cudaStreamSynchronize(h->stream));
    for (int c = 0; c < chunks; c++) x.chunk_row[c] = (int)(tp[c] & ROW_MASK);
    for (int c = 1; c < chunks; c++)
        CUX(cudaMemcpyAsync(&rp[c], pl.row_ptr + x.chunk_row[c], sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUX(cudaStreamSynchronize(h->stream));
    const long long tile_nnz = (long long)OMEGA * pl.sigma;
    x.n_boundary_rows = 0;
    for (int c = 1; c < chunks; c++) {
        // row chunk_row[c] holds the first non-zero of block c: it crosses in unless it starts exactly there
        x.chunk_carried[c] = (long long)rp[c] != x.chunk_tile[c] * tile_nnz;
        if (x.chunk_tile[c] == x.chunk_tile[c + 1] && c + 1 < chunks) x.chunk_carried[c] = 0;   // empty block
        if (x.chunk_carried[c] &&
            (x.n_boundary_rows == 0 || x.boundary_rows[x.n_boundary_rows - 1] != x.chunk_row[c]))
            x.boundary_rows[x.n_boundary_rows++] = x.chunk_row[c];
    }
    x.chunk_row[0] = 0;   // leading empty rows belong to the first block (cleared by the prologue)
    x.table = ChunkTable();
    x.table.n = chunks;
    for (int c = 0; c <= chunks; c++) x.table.tile_begin[c] = x.chunk_tile[c];
    for (int c = 0; c < chunks; c++) x.table.skip_row[c] = x.chunk_carried[c] ? x.chunk_row[c] : -1;
    // Order: ship-heavy blocks first (Johnson's rule for a two-stage pipeline whose first stage -- the SpMV of a block --
    // costs about the same for every block); uniform matrices keep the natural order.
    x.chunk_order.resize(chunks);
    std::iota(x.chunk_order.begin(), x.chunk_order.end(), 0);
    auto rows_of = [&](int c) { return (long long)x.chunk_row[c + 1] - x.chunk_row[c]; };
    long long lo = rows_of(0), hi = rows_of(0);
    for (int c = 1; c < chunks; c++) { lo = std::min(lo, rows_of(c)); hi = std::max(hi, rows_of(c)); }
    x.reordered = hi > 2 * lo + 1024;
    if (x.reordered)
        std::stable_sort(x.chunk_order.begin(), x.chunk_order.end(), [&](int a, int b) { return rows_of(a) > rows_of(b); });
    x.chunks = chunks;
    return CSR5B200_SUCCESS;
}

cudaError_t spmv_part(csr5b200_handle_t h, double alpha, double beta, void *y, const ShardCtx *sh, const SpmvCall &call,
                      cudaStream_t stream)
{
    if (h->pl.value_bytes == 8)
        return launch_spmv_part_f64(h->pl, h->tune, alpha, beta, static_cast<double *>(y), sh, call, stream,
                                    &h->kernel_in_use, &h->launches_per_spmv);
    return launch_spmv_part_f32(h->pl, h->tune, (float)alpha, (float)beta, static_cast<float *>(y), sh, call, stream,
                                &h->kernel_in_use, &h->launches_per_spmv);
}

}  // namespace

}  // namespace csr5

using namespace csr5;

extern "C" {

int csr5b200_spmv_allgather(csr5b200_handle_t h, double alpha, double beta, const csr5b200_exchange *ex)
{
    if (!h || !ex) return CSR5B200_INVALID_ARGUMENT;
    if (h->format == CSR5B200_FORMAT_CSR) return CSR5B200_UNSUPPORTED_CSR_SPMV;
    if (h->format != CSR5B200_FORMAT_CSR5) return CSR5B200_UNKNOWN_FORMAT;
    const Plan &pl = h->pl;
    const int world = ex->world, rank = ex->rank;
    if (world < 1 || world > CSR5B200_MAX_SCATTER || rank < 0 || rank >= world || ex->row_begin < 0)
        return CSR5B200_INVALID_ARGUMENT;
    if (!pl.x && pl.nnz > 0) return CSR5B200_INVALID_ARGUMENT;
    for (int k = 0; k < world; k++)
        if (!ex->y_full[k]) return CSR5B200_INVALID_ARGUMENT;
    bool barriers = world > 1;
    for (int k = 0; k < world; k++)
        if (!ex->flags[k]) barriers = false;
    int transport = ex->transport;
    if (transport < 0 || transport > CSR5B200_TRANSPORT_NONE) return CSR5B200_INVALID_ARGUMENT;
    if (transport == CSR5B200_TRANSPORT_AUTO) {
        // measured on 2, 4 and 8 B200 (profiles/r02_bench_c2_n{2,4,8}_sweep*.json).  Matrices whose tiles store runs
        // of consecutive rows (no empty rows, short rows): up to 4 GPUs the SpMV kernel's own peer stores win (C2:
        // 0.337 / 0.509 ms at 2 / 4 GPUs against 0.430 / 0.554 through the multicast address), from 5 up the NVSwitch
        // multicast address (8 GPUs: 0.966 against 1.210) -- one store per row instead of N - 1.  Everything else
        // (scattered row stores, shards with very different row counts): multicast from 3 GPUs up (push grid without
        // a multicast address), the copy engine between 2.
        const bool consecutive = !pl.needs_zero_fill && pl.m > 0 && (long long)pl.nnz / pl.m <= 64 && beta == 0.0;
        if (consecutive && world <= 4) transport = CSR5B200_TRANSPORT_IN_KERNEL;
        else if (world >= 3) transport = ex->y_multicast ? CSR5B200_TRANSPORT_SM_MULTICAST : CSR5B200_TRANSPORT_SM_PUSH;
        else transport = CSR5B200_TRANSPORT_COPY_ENGINE;
    }
    if (transport == CSR5B200_TRANSPORT_SM_MULTICAST && !ex->y_multicast) return CSR5B200_INVALID_ARGUMENT;
    if (world == 1) transport = CSR5B200_TRANSPORT_NONE;
    if (transport == CSR5B200_TRANSPORT_IN_KERNEL && beta != 0.0) return CSR5B200_INVALID_ARGUMENT;
    if (h->ignore_alpha) alpha = 1.0;
    int err = ensure_state(h);
    if (err) return err;
    ExchangeState &x = h->ex;
    const size_t vb = (size_t)pl.value_bytes;
    char *y_local = static_cast<char *>(ex->y_full[rank]) + (size_t)ex->row_begin * vb;
    cudaStream_t S = h->stream;
    h->launches_per_spmv = 0;
    h->tune.ev_begin = h->tune.ev_end = nullptr;

    if (pl.m <= 0 || pl.p == 0) {   // nothing to compute: clear / scale the rows, then only the barriers
        SpmvCall all;
        CUX(spmv_part(h, alpha, beta, y_local, nullptr, all, S));
        if (barriers && ex->entry_barrier) {
            CUX(launch_flag_barrier(ex->flags, rank, world, 0, x.epoch, x.status, ex->timeout_ms, S));
            ++h->launches_per_spmv;
        }
        // a shard of empty rows only: its (cleared / scaled) rows still have to reach every peer
        if (pl.m > 0 && transport != CSR5B200_TRANSPORT_NONE)
            for (int k = 0; k < world; k++)
                if (k != rank)
                    CUX(cudaMemcpyAsync(static_cast<char *>(ex->y_full[k]) + (size_t)ex->row_begin * vb, y_local,
                                        (size_t)pl.m * vb, cudaMemcpyDefault, S));
        if (barriers) {
            CUX(launch_flag_barrier(ex->flags, rank, world, 1, x.epoch, x.status, ex->timeout_ms, S));
            ++h->launches_per_spmv;
        }
        x.last_transport = transport;
        x.last_chunks = 0;
        return CSR5B200_SUCCESS;
    }

    // Default block count, measured on 8 GPUs (profiles/r02_bench_c2_n8_sweep2.json, r02_bench_c5_n8_chunks_trace.json):
    // 12 for shards whose blocks carry similar numbers of rows (C2: 6 / 8 / 12 / 16 blocks -> 0.994 / 0.985 / 0.966 /
    // 0.965 ms), 8 when they differ so much that the blocks are reordered (R-MAT 25: 4 / 8 / 12 / 16 / 24 blocks ->
    // 0.581 / 0.575 / 0.591 / 0.609 / 0.668 ms: every block costs a push launch on the critical path).
    int chunks = ex->chunks > 0 ? ex->chunks : (x.auto_chunks > 0 ? x.auto_chunks : 12);
    if (transport == CSR5B200_TRANSPORT_IN_KERNEL) chunks = 1;
    if ((err = ensure_chunks(h, chunks))) return err;
    if (ex->chunks <= 0 && x.auto_chunks == 0 && transport != CSR5B200_TRANSPORT_IN_KERNEL) {
        x.auto_chunks = x.reordered ? 8 : 12;
        if (x.auto_chunks != x.chunks && (err = ensure_chunks(h, x.auto_chunks))) return err;
    }
    chunks = x.chunks;
    const int push_ctas = ex->push_ctas > 0 ? ex->push_ctas : 48;

    if (!x.warmed) {
        // Load every module of the step now: CUDA loads kernels lazily at their first launch, and that load
        // synchronises the context -- against a barrier kernel that is spinning for a peer it would deadlock.
        // One throw-away SpMV into a scratch vector (y itself must stay what the caller handed in when beta != 0).
        void *scratch = nullptr;
        CUX(cudaMalloc(&scratch, (size_t)pl.m * vb));
        SpmvCall all;
        cudaError_t e = spmv_part(h, alpha, beta, scratch, nullptr, all, S);
        void *none[CSR5B200_MAX_SCATTER] = {};
        if (e == cudaSuccess) e = launch_push_rows((int)vb, scratch, none, 0, 0, pl.m < 1024 ? pl.m : 1024, 1, 256, S);
        if (e == cudaSuccess) e = launch_flag_barrier(nullptr, 0, 0, 0, x.epoch, x.status, 0, S);
        if (e == cudaSuccess) e = cudaStreamSynchronize(S);
        cudaFree(scratch);
        if (e != cudaSuccess) return handle_cuda_fail(h, e);
        x.warmed = true;
        h->launches_per_spmv = 0;
    }

    // ---- legacy scheme inside the new step: the SpMV kernel stores to every destination itself ----------------
    if (transport == CSR5B200_TRANSPORT_IN_KERNEL) {
        void *dst[CSR5B200_MAX_SCATTER];
        for (int k = 0; k < world; k++) dst[k] = static_cast<char *>(ex->y_full[k]) + (size_t)ex->row_begin * vb;
        ShardCtx sh;
        sh.n_dst = world;
        sh.y_dst = dst;
        sh.multicast = 0;
        sh.exchange = 1;
        if (barriers && ex->entry_barrier) {
            CUX(launch_flag_barrier(ex->flags, rank, world, 0, x.epoch, x.status, ex->timeout_ms, S));
            ++h->launches_per_spmv;
        }
        SpmvCall all;
        CUX(spmv_part(h, alpha, 0.0, y_local, &sh, all, S));
        if (barriers) {
            CUX(launch_flag_barrier(ex->flags, rank, world, 1, x.epoch, x.status, ex->timeout_ms, S));
            ++h->launches_per_spmv;
        }
        x.last_transport = transport;
        x.last_chunks = 1;
        return CSR5B200_SUCCESS;
    }

    // ---- prologue on S, then fork ---------------------------------------------------------------------------------
    {
        SpmvCall pro;
        pro.tiles = pro.calibrate = false;
        CUX(spmv_part(h, alpha, beta, y_local, nullptr, pro, S));
    }
    const bool ship = transport != CSR5B200_TRANSPORT_NONE;
    const bool by_ce = transport == CSR5B200_TRANSPORT_COPY_ENGINE;
    const bool trace = x.trace;
    if (trace) {
        if (!x.tv_begin) {
            CUX(cudaEventCreate(&x.tv_begin));
            CUX(cudaEventCreate(&x.tv_end));
            for (int i = 0; i < MAX_CHUNKS; i++) {
                CUX(cudaEventCreate(&x.tv_tiles[i]));
                CUX(cudaEventCreate(&x.tv_cal[i]));
                CUX(cudaEventCreate(&x.tv_ship[i]));
            }
        }
        x.traced_chunks = chunks;
        CUX(cudaEventRecord(x.tv_begin, S));
    }
    const bool by_mc = transport == CSR5B200_TRANSPORT_SM_MULTICAST;
    CUX(cudaEventRecord(x.ev_begin, S));
    CUX(cudaStreamWaitEvent(x.side, x.ev_begin, 0));
    if (chunks > 1) CUX(cudaStreamWaitEvent(x.work1, x.ev_begin, 0));
    if (barriers && ex->entry_barrier && ship) {
        CUX(launch_flag_barrier(ex->flags, rank, world, 0, x.epoch, x.status, ex->timeout_ms, x.side));
        ++h->launches_per_spmv;
    }

    // destinations of the SM transports, relative to this shard's first row
    void *dst0[CSR5B200_MAX_SCATTER] = {};
    int n_dst = 0;
    const size_t seg = (size_t)ex->row_begin * vb;
    if (by_mc) {
        dst0[n_dst++] = static_cast<char *>(ex->y_multicast) + seg;
    } else {
        for (int k = 0; k < world; k++)
            if (k != rank) dst0[n_dst++] = static_cast<char *>(ex->y_full[k]) + seg;
    }

    for (int i = 0; i < chunks; i++) {
        const int c = x.chunk_order[i];
        cudaStream_t W = (i & 1) ? x.work1 : S;
        SpmvCall blk;
        blk.tile_begin = x.chunk_tile[c];
        blk.tile_end = x.chunk_tile[c + 1];
        blk.tail = c == chunks - 1;
        blk.prologue = false;
        blk.calibrate = false;
        CUX(spmv_part(h, alpha, beta, y_local, nullptr, blk, W));
        CUX(cudaEventRecord(x.ev_chunk[i], W));
        if (trace) CUX(cudaEventRecord(x.tv_tiles[i], W));

        // carry pass of the block (all but the carries into the row that crossed in), then its rows leave
        CUX(cudaStreamWaitEvent(x.side, x.ev_chunk[i], 0));
        SpmvCall cal = blk;
        cal.tiles = false;
        cal.calibrate = true;
        cal.skip_row = x.table.skip_row[c];
        CUX(spmv_part(h, alpha, beta, y_local, nullptr, cal, x.side));
        if (trace) {
            CUX(cudaEventRecord(x.tv_cal[i], x.side));
            x.traced_ship[i] = false;
        }
        const long long ra = (long long)x.chunk_row[c] + (x.chunk_carried[c] ? 1 : 0), rb = x.chunk_row[c + 1];
        if (!ship || rb <= ra) continue;
        CUX(cudaEventRecord(x.ev_cal[i], x.side));
        const char *src = y_local + (size_t)ra * vb;
        if (by_ce) {
            for (int k = 0; k < world; k++) {
                if (k == rank) continue;
                CUX(cudaStreamWaitEvent(x.ce[k], x.ev_cal[i], 0));
                CUX(cudaMemcpyAsync(static_cast<char *>(ex->y_full[k]) + seg + (size_t)ra * vb, src, (size_t)(rb - ra) * vb,
                                    cudaMemcpyDefault, x.ce[k]));
                ++h->launches_per_spmv;
            }
        } else {
            void *dst[CSR5B200_MAX_SCATTER] = {};
            for (int k = 0; k < n_dst; k++) dst[k] = static_cast<char *>(dst0[k]) + (size_t)ra * vb;
            CUX(cudaStreamWaitEvent(x.ship, x.ev_cal[i], 0));
            CUX(launch_push_rows((int)vb, src, dst, n_dst, by_mc ? 1 : 0, rb - ra, push_ctas, ex->push_threads, x.ship));
            ++h->launches_per_spmv;
            if (trace) {
                CUX(cudaEventRecord(x.tv_ship[i], x.ship));
                x.traced_ship[i] = true;
            }
        }
    }

    // ---- join: boundary pass, crossing rows, barrier -----------------------------------------------------------
    if (chunks > 1) {
        CUX(cudaEventRecord(x.ev_work1_done, x.work1));
        CUX(cudaStreamWaitEvent(S, x.ev_work1_done, 0));
    }
    CUX(cudaEventRecord(x.ev_side_done, x.side));
    CUX(cudaStreamWaitEvent(S, x.ev_side_done, 0));
    if (x.n_boundary_rows > 0) {
        SpmvCall bnd;
        bnd.boundary = &x.table;
        CUX(spmv_part(h, alpha, beta, y_local, nullptr, bnd, S));
        if (ship) {
            CUX(launch_push_row_list((int)vb, y_local, dst0, n_dst, by_mc ? 1 : 0, x.boundary_rows, x.n_boundary_rows, S));
            ++h->launches_per_spmv;
        }
    }
    if (ship && by_ce) {
        for (int k = 0; k < world; k++) {
            if (k == rank) continue;
            CUX(cudaEventRecord(x.ev_ce_done[k], x.ce[k]));
            CUX(cudaStreamWaitEvent(S, x.ev_ce_done[k], 0));
        }
    } else if (ship) {
        CUX(cudaEventRecord(x.ev_ship_done, x.ship));
        CUX(cudaStreamWaitEvent(S, x.ev_ship_done, 0));
    }
    if (barriers) {
        CUX(launch_flag_barrier(ex->flags, rank, world, 1, x.epoch, x.status, ex->timeout_ms, S));
        ++h->launches_per_spmv;
    }
    if (trace) CUX(cudaEventRecord(x.tv_end, S));
    x.last_transport = transport;
    x.last_chunks = chunks;
    return CSR5B200_SUCCESS;
}

int csr5b200_exchange_trace(csr5b200_handle_t h, float *ms, int capacity, int *count)
{
    if (!h || !count || (capacity > 0 && !ms)) return CSR5B200_INVALID_ARGUMENT;
    *count = 0;
    ExchangeState &x = h->ex;
    if (!x.trace || !x.tv_begin || x.traced_chunks <= 0) return CSR5B200_SUCCESS;
    CUX(cudaStreamSynchronize(h->stream));
    auto put = [&](cudaEvent_t e, bool valid) {
        float v = -1.f;
        if (valid && cudaEventElapsedTime(&v, x.tv_begin, e) != cudaSuccess) { v = -1.f; cudaGetLastError(); }
        if (*count < capacity) ms[(*count)++] = v;
    };
    for (int i = 0; i < x.traced_chunks; i++) {
        put(x.tv_tiles[i], true);
        put(x.tv_cal[i], true);
        put(x.tv_ship[i], x.traced_ship[i]);
    }
    put(x.tv_end, true);
    return CSR5B200_SUCCESS;
}

int csr5b200_exchange_status(csr5b200_handle_t h)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    CUX(cudaStreamSynchronize(h->stream));
    if (!h->ex.ready) return CSR5B200_SUCCESS;
    int st = 0;
    CUX(cudaMemcpy(&st, h->ex.status, sizeof(int), cudaMemcpyDeviceToHost));
    if (st) {
        CUX(cudaMemset(h->ex.status, 0, sizeof(int)));
        return CSR5B200_EXCHANGE_TIMEOUT;
    }
    return CSR5B200_SUCCESS;
}

}  // extern "C"
