// csr5_capi.cu -- handle state machine and the extern "C" ABI of libcsr5_b200.so
// (include/csr5_b200.h).  Mirrors anonymouslibHandle<int, unsigned int, VT>
// (CSR5_cuda/anonymouslib_cuda.h:11-318): CSR arrays are borrowed, the five CSR5 arrays are owned,
// col/val are permuted in place between asCSR5() and asCSR().
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "csr5_handle.h"

using namespace csr5;

namespace csr5 {
int handle_cuda_fail(csr5b200_handle_t h, cudaError_t e)
{
    if (h) h->last_cuda_error = (int)e;
    return CSR5B200_CUDA_ERROR;
}
}  // namespace csr5

namespace {

int cuda_fail(csr5b200_handle_t h, cudaError_t e) { return handle_cuda_fail(h, e); }

#define CU(h, call)                                        \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return cuda_fail(h, e__);  \
    } while (0)

// The CSR5 arrays go back to the handle's buffer pool (the memory is kept for the next asCSR5()).
void release_csr5_arrays(csr5b200_handle_t h)
{
    Plan &pl = h->pl;
    cudaFree(pl.hot_col);
    cudaFree(pl.hot_x);
    pl.hot_col = nullptr;
    pl.hot_x = nullptr;
    pl.hot_k = 0;
    pl.hot_coverage = 0.0;
    pl.tile_ptr = nullptr;
    pl.desc = nullptr;
    pl.desc_off_ptr = nullptr;
    pl.desc_off = nullptr;
    pl.calibrator = nullptr;
    pl.dev_flags = nullptr;
    h->ex.chunks = 0;   // row-block boundaries of the exchange belong to the released tile_ptr
    h->ex.auto_chunks = 0;
    h->ex.chunk_tile.clear();
    h->ex.chunk_row.clear();
}

void free_pool(csr5b200_handle_t h)
{
    for (int k = 0; k < csr5b200_handle_s::POOL_SLOTS; k++) {
        cudaFree(h->pool[k]);
        h->pool[k] = nullptr;
        h->pool_cap[k] = 0;
    }
    for (auto &e : h->ev_conv) {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
}

// A buffer of at least `bytes` from slot `slot` of the pool (grown when too small).
cudaError_t pool_get(csr5b200_handle_t h, int slot, size_t bytes, void **out)
{
    if (bytes == 0) bytes = 16;
    if (h->pool_cap[slot] < bytes) {
        cudaFree(h->pool[slot]);
        h->pool[slot] = nullptr;
        h->pool_cap[slot] = 0;
        const cudaError_t e = cudaMalloc(&h->pool[slot], bytes);
        if (e != cudaSuccess) return e;
        h->pool_cap[slot] = bytes;
    }
    *out = h->pool[slot];
    return cudaSuccess;
}

constexpr size_t MAX_TIMED_SPMV = 4096;

// anonymouslib_cuda.h:294-318 -- r / s / t / u = 4 / 32 / 256 / 6 on k = nnz / m
int auto_sigma(int m, int nnz)
{
    const int k = m > 0 ? nnz / m : 0;
    if (k <= 4) return 4;
    if (k <= 32) return k;
    if (k <= 256) return 32;
    return 6;
}

// CSR5B200_OPT_SIGMA_RULE = 1: the table measured on B200 (profiles/r02_sigma_rule.md; sweep of nnz/row 2..300 x
// {FP64, FP32} x sigma 4..32, tools/gpu/sweep_sigma.py), replacing the reference's Maxwell-era one above.  Same input
// (k = nnz / m) so that it stays a drop-in for setSigma(AUTO).  What the sweep shows and the rule keeps:
//   * short rows: sigma = k makes tiles of a few hundred bytes whose descriptor / tile_ptr / launch overhead costs
//     8-16 % (FP64) and 16-29 % (FP32) against sigma 12-16 -> a floor on sigma;
//   * k > 256: the reference falls back to sigma = 6 (its `u`), 7 % (FP64) to 33 % (FP32) slower than 32;
//   * 12 <= k <= 256: every sigma from 12 up is within 1-4 % and the ranking changes from box to box and with the
//     matrix size (profiles/r02_ab_parked_row_stores.txt: sigma 16 beats 14 by 4 % on C2 where the sweep's 4 M-row
//     matrix had it 2 % behind) -> the reference's choice is kept there.
int auto_sigma_b200(int m, int nnz, int value_bytes)
{
    const int k = m > 0 ? nnz / m : 0;
    if (value_bytes == 8) {
        if (k <= 2) return 4;
        if (k <= 5) return 12;
        if (k <= 11) return 14;
    } else {
        if (k <= 2) return 8;
        if (k == 3) return 12;
        if (k <= 11) return 16;
    }
    return k <= 32 ? k : 32;
}

// Hot-column table: pick the most referenced columns of the CSR5 tiles, at most
// `capacity`, by bisecting the reference-count threshold; tag their occurrences in col (bit 31 | slot).
// Auto mode keeps the table only when it would serve >= 25 % of the tiles' x references.
int build_hot_table(csr5b200_handle_t h)
{
    Plan &pl = h->pl;
    const int mode = h->tune.hot_columns;
    if (mode == 0 || pl.p < 2 || pl.n <= 0) return CSR5B200_SUCCESS;
    const size_t vb = (size_t)pl.value_bytes;
    int capacity = mode > 0 ? mode : (int)((128 * 1024) / vb);   // auto: 128 KB of shared memory
    const int cap_max = (int)((200 * 1024) / vb);
    if (capacity > cap_max) capacity = cap_max;
    const long long limit = (long long)(pl.p - 1) * OMEGA * pl.sigma;
    int *cnt = nullptr, *slot = nullptr;
    unsigned long long *out2 = nullptr;
    void *scratch = nullptr;
    const size_t scratch_bytes = scan_scratch_bytes(pl.n);
    auto done = [&](int code) {
        cudaFree(cnt); cudaFree(slot); cudaFree(out2); cudaFree(scratch);
        return code;
    };
#define CUH(call)                                                   \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return done(cuda_fail(h, e__));     \
    } while (0)
    CUH(cudaMalloc(&cnt, (size_t)pl.n * sizeof(int)));
    CUH(cudaMalloc(&out2, 2 * sizeof(unsigned long long)));
    CUH(cudaMemsetAsync(cnt, 0, (size_t)pl.n * sizeof(int), h->stream));
    CUH(launch_hot_count(pl.col, limit, cnt, h->tune.num_sms, h->stream));
    // smallest threshold whose column set fits the table (set size is non-increasing in the threshold): from the
    // histogram of the counts in ONE pass and one read-back; counts beyond the last bin fall back to bisection
    unsigned long long res[2] = {0, 0};
    auto count_ge = [&](long long thr) -> cudaError_t {
        cudaError_t e = launch_hot_count_ge(cnt, pl.n, (int)thr, out2, h->tune.num_sms, h->stream);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(res, out2, sizeof(res), cudaMemcpyDeviceToHost, h->stream);
        if (e != cudaSuccess) return e;
        return cudaStreamSynchronize(h->stream);
    };
    long long lo = 1, hi = limit + 1;   // invariant: set(hi) fits; set(lo - 1) does not (or lo == 1)
    {
        const int bins = hot_hist_bins();
        unsigned int *d_cols = nullptr;
        unsigned long long *d_refs = nullptr;
        std::vector<unsigned int> hc(bins);
        std::vector<unsigned long long> hr(bins);
        cudaError_t e = cudaMalloc(&d_cols, bins * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMalloc(&d_refs, bins * sizeof(unsigned long long));
        if (e == cudaSuccess) e = launch_hot_hist(cnt, pl.n, d_cols, d_refs, h->tune.num_sms, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hc.data(), d_cols, bins * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hr.data(), d_refs, bins * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(d_cols);
        cudaFree(d_refs);
        if (e != cudaSuccess) return done(cuda_fail(h, e));
        unsigned long long cols = 0, refs = 0;
        int t = bins - 1;
        if (hc[t] > (unsigned int)capacity) {
            lo = bins - 1;                       // even the clamped top bin overflows the table: bisect above it
        } else {
            for (; t >= 1; t--) {
                if (cols + hc[t] > (unsigned long long)capacity) break;
                cols += hc[t];
                refs += hr[t];
            }
            lo = hi = t + 1;                     // threshold t + 1: everything counted so far
            res[0] = cols;
            res[1] = refs;
        }
    }
    while (lo < hi) {
        const long long mid = lo + (hi - lo) / 2;
        CUH(count_ge(mid));
        if (res[0] <= (unsigned long long)capacity) hi = mid; else lo = mid + 1;
    }
    if (res[0] == 0 || lo >= hot_hist_bins() - 1) CUH(count_ge(lo));
    const int threshold = (int)lo;
    const int hot_k = (int)res[0];
    const double coverage = limit > 0 ? (double)res[1] / (double)limit : 0.0;
    if (hot_k <= 0 || (mode < 0 && coverage < 0.25)) return done(CSR5B200_SUCCESS);

    CUH(cudaMalloc(&slot, (size_t)(pl.n + 1) * sizeof(int)));
    CUH(cudaMalloc(&scratch, scratch_bytes));
    CUH(cudaMalloc(&pl.hot_col, (size_t)hot_k * sizeof(int)));
    CUH(cudaMalloc(&pl.hot_x, ((size_t)hot_k * vb + 15) / 16 * 16));
    CUH(cudaMemsetAsync(pl.hot_x, 0, ((size_t)hot_k * vb + 15) / 16 * 16, h->stream));
    CUH(launch_hot_flags(cnt, pl.n, threshold, slot, h->stream));
    CUH(launch_exclusive_scan(slot, pl.n + 1, scratch, scratch_bytes, h->stream));
    CUH(launch_hot_assign(cnt, pl.n, threshold, slot, pl.hot_col, pl.col, limit, h->tune.num_sms, h->stream));
    CUH(cudaStreamSynchronize(h->stream));
#undef CUH
    pl.hot_k = hot_k;
    pl.hot_coverage = coverage;
    return done(CSR5B200_SUCCESS);
}

}  // namespace

extern "C" {

int csr5b200_create(int m, int n, int value_bytes, csr5b200_handle_t *out)
{
    if (!out || m < 0 || n < 0) return CSR5B200_INVALID_ARGUMENT;
    if (value_bytes != 4 && value_bytes != 8) return CSR5B200_UNSUPPORTED_VALUE_TYPE;
    csr5b200_handle_t h = new (std::nothrow) csr5b200_handle_s();
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    h->pl.m = m;
    h->pl.n = n;
    h->pl.value_bytes = value_bytes;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
        h->tune.num_sms = sms;
    // opt-in for unmodified callers of the reference API (the shim only ever calls setSigma(AUTO))
    const char *rule = std::getenv("CSR5B200_SIGMA_RULE");
    if (rule && (!std::strcmp(rule, "b200") || !std::strcmp(rule, "1"))) h->sigma_rule = 1;
    *out = h;
    return CSR5B200_SUCCESS;
}

int csr5b200_warmup(csr5b200_handle_t h)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    CU(h, launch_warmup(h->stream));
    return CSR5B200_SUCCESS;
}

int csr5b200_input_csr(csr5b200_handle_t h, int nnz, int *row_ptr, int *col, void *val)
{
    if (!h || nnz < 0) return CSR5B200_INVALID_ARGUMENT;
    if (h->format == CSR5B200_FORMAT_CSR5) {  // the reference would leak; restore first
        const int err = csr5b200_as_csr(h);
        if (err) return err;
    }
    h->format = CSR5B200_FORMAT_CSR;
    h->pl.nnz = nnz;
    h->pl.row_ptr = row_ptr;
    h->pl.col = col;
    h->pl.val = val;
    return CSR5B200_SUCCESS;
}

int csr5b200_set_x(csr5b200_handle_t h, void *x)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    h->pl.x = x;  // no texture object: sm_100a gathers through the read-only LDG path
    return CSR5B200_SUCCESS;
}

int csr5b200_set_sigma(csr5b200_handle_t h, int sigma)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    if (sigma == CSR5B200_AUTO_TUNED_SIGMA)
        sigma = h->sigma_rule == 1 ? auto_sigma_b200(h->pl.m, h->pl.nnz, h->pl.value_bytes) : auto_sigma(h->pl.m, h->pl.nnz);
    h->pl.sigma = sigma;
    return CSR5B200_SUCCESS;
}

int csr5b200_set_stream(csr5b200_handle_t h, void *cuda_stream)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return CSR5B200_SUCCESS;
}

int csr5b200_set_option(csr5b200_handle_t h, int option, int value)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    switch (option) {
        case CSR5B200_OPT_KERNEL:
            if (value < 0 || value > 4 || value == 3) return CSR5B200_INVALID_ARGUMENT;
            h->tune.kernel = value;
            break;
        case CSR5B200_OPT_IGNORE_ALPHA: h->ignore_alpha = value != 0; break;
        case CSR5B200_OPT_TMA_STAGES: h->tune.tma_stages = value; break;
        case CSR5B200_OPT_TMA_WARPS: h->tune.tma_warps = value; break;
        case CSR5B200_OPT_CTAS_PER_SM: h->tune.ctas_per_sm = value; break;
        case CSR5B200_OPT_KERNEL_TIMING: h->kernel_timing = value != 0; break;
        case CSR5B200_OPT_DIRECT_WPB: h->tune.direct_wpb = value; break;
        case CSR5B200_OPT_DIRECT_NCH: h->tune.direct_nch = value; break;
        case CSR5B200_OPT_HOT_COLUMNS: h->tune.hot_columns = value; break;
        case CSR5B200_OPT_HOT_THREADS: h->tune.hot_threads = value; break;
        case CSR5B200_OPT_EXCHANGE_TRACE: h->ex.trace = value != 0; break;
        case CSR5B200_OPT_DETERMINISTIC: h->tune.deterministic = value != 0; break;
        case CSR5B200_OPT_SIGMA_RULE:
            if (value < 0 || value > 1) return CSR5B200_INVALID_ARGUMENT;
            h->sigma_rule = value;
            break;
        case CSR5B200_OPT_EXCHANGE:
            if (value < 0 || value > 2) return CSR5B200_INVALID_ARGUMENT;
            h->tune.exchange = value;
            break;
        default: return CSR5B200_INVALID_ARGUMENT;
    }
    return CSR5B200_SUCCESS;
}

int csr5b200_as_csr5(csr5b200_handle_t h)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    if (h->format == CSR5B200_FORMAT_CSR5) return CSR5B200_SUCCESS;
    if (h->format != CSR5B200_FORMAT_CSR) return CSR5B200_UNKNOWN_FORMAT;
    Plan &pl = h->pl;
    if (pl.sigma < SIGMA_MIN || pl.sigma > SIGMA_MAX) return CSR5B200_CSR_TO_CSR5_FAILED;

    // anonymouslib_cuda.h:121-137
    int base = 2;
    pl.bit_y = 1;
    while (base < OMEGA * pl.sigma) { base *= 2; pl.bit_y++; }
    base = 2;
    pl.bit_ss = 1;
    while (base < OMEGA) { base *= 2; pl.bit_ss++; }
    if (pl.bit_y + pl.bit_ss > 31) return CSR5B200_UNSUPPORTED_CSR5_OMEGA;
    pl.num_packet = (pl.bit_y + pl.bit_ss + pl.sigma + 31) / 32;
    const long long tile = (long long)OMEGA * pl.sigma;
    pl.p = (int)((pl.nnz + tile - 1) / tile);
    pl.tail_start = 0;
    pl.num_offsets = 0;
    pl.needs_zero_fill = 0;
    pl.has_carries = 1;

    if (pl.p == 0) {  // empty matrix: spmv() only clears y
        h->format = CSR5B200_FORMAT_CSR5;
        return CSR5B200_SUCCESS;
    }

    const size_t vb = (size_t)pl.value_bytes;
    void *scan_scratch = nullptr;
    const size_t scan_bytes = scan_scratch_bytes(pl.p);
    bool transposed = false;
    auto fail = [&](cudaError_t e) {
        // leave the caller's arrays as they were handed in: a failure after the in-place transpose was enqueued
        // undoes it (the transpose is its own inverse pair r2c / c2r)
        if (transposed && launch_transpose(pl, false, h->stream) == cudaSuccess) cudaStreamSynchronize(h->stream);
        release_csr5_arrays(h);
        return cuda_fail(h, e);
    };
#define CUF(call)                                   \
    do {                                            \
        cudaError_t e__ = (call);                   \
        if (e__ != cudaSuccess) return fail(e__);   \
    } while (0)

    const auto host_t0 = std::chrono::steady_clock::now();
    typedef csr5b200_handle_s HS;
    CUF(pool_get(h, HS::POOL_TILE_PTR, (size_t)(pl.p + 1) * sizeof(uint32_t), reinterpret_cast<void **>(&pl.tile_ptr)));
    CUF(pool_get(h, HS::POOL_DESC, (size_t)pl.p * OMEGA * pl.num_packet * sizeof(uint32_t), reinterpret_cast<void **>(&pl.desc)));
    CUF(pool_get(h, HS::POOL_DESC_OFF_PTR, (size_t)(pl.p + 1) * sizeof(int), reinterpret_cast<void **>(&pl.desc_off_ptr)));
    CUF(pool_get(h, HS::POOL_CALIBRATOR, (size_t)pl.p * vb, &pl.calibrator));
    CUF(pool_get(h, HS::POOL_FLAGS, 8 * sizeof(int), reinterpret_cast<void **>(&pl.dev_flags)));
    CUF(pool_get(h, HS::POOL_SCAN, scan_bytes, &scan_scratch));
    for (auto &e : h->ev_conv)
        if (!e) CUF(cudaEventCreate(&e));
    h->convert_alloc_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    CUF(cudaMemsetAsync(pl.dev_flags, 0, 8 * sizeof(int), h->stream));
    CUF(cudaMemsetAsync(pl.calibrator, 0, (size_t)pl.p * vb, h->stream));

    CUF(cudaEventRecord(h->ev_conv[0], h->stream));
    CUF(launch_tile_ptr(pl, h->stream));
    CUF(cudaEventRecord(h->ev_conv[1], h->stream));
    CUF(launch_tile_desc(pl, h->stream));
    CUF(cudaEventRecord(h->ev_conv[2], h->stream));
    CUF(launch_scan_offsets(pl, scan_scratch, scan_bytes, h->stream));
    CUF(cudaEventRecord(h->ev_conv[3], h->stream));

    // One blocking read-back (the reference does three, anonymouslib_cuda.h:166, format_cuda.h:331,342):
    // tail start, first tile's row, number of empty-row table entries, "any dirty tile" flag.
    uint32_t tp_last = 0, tp_first = 0;
    int num_offsets = 0, any_dirty = 0, any_carry = 1;
    CUF(cudaMemcpyAsync(&tp_last, pl.tile_ptr + pl.p - 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CUF(cudaMemcpyAsync(&tp_first, pl.tile_ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CUF(cudaMemcpyAsync(&num_offsets, pl.desc_off_ptr + pl.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUF(cudaMemcpyAsync(&any_dirty, pl.dev_flags, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUF(cudaMemcpyAsync(&any_carry, pl.dev_flags + 1, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUF(cudaStreamSynchronize(h->stream));
    // no tile continues a row (e.g. 16 nnz/row at sigma 16): spmv() skips the carry pass, one launch per SpMV
    pl.has_carries = any_carry ? 1 : 0;
    pl.tail_start = (int)(tp_last & ROW_MASK);
    pl.num_offsets = num_offsets;
    // Rows that no tile stores: empty rows inside CSR5 tiles and the empty rows in front of the
    // first non-empty row.  (Empty rows of the tail are written by the tail warps.)
    pl.needs_zero_fill = (any_dirty || (tp_first & ROW_MASK) > 0) ? 1 : 0;

    CUF(cudaEventRecord(h->ev_conv[4], h->stream));
    if (num_offsets > 0) {
        CUF(pool_get(h, HS::POOL_DESC_OFF, (size_t)num_offsets * sizeof(int), reinterpret_cast<void **>(&pl.desc_off)));
        CUF(launch_desc_offset(pl, h->stream));
    }
    CUF(cudaEventRecord(h->ev_conv[5], h->stream));
    transposed = true;
    CUF(launch_transpose(pl, true, h->stream));
    CUF(cudaEventRecord(h->ev_conv[6], h->stream));
    CUF(cudaStreamSynchronize(h->stream));
    {
        // device time of the phases: tile_ptr, tile_desc, scan, (read-back), desc_offset, transpose
        static const int from[5] = {0, 1, 2, 4, 5};
        for (int k = 0; k < 5; k++)
            if (cudaEventElapsedTime(&h->convert_ms[k], h->ev_conv[from[k]], h->ev_conv[from[k] + 1]) != cudaSuccess)
                h->convert_ms[k] = -1.f;
        cudaGetLastError();
    }
    {
        const int err = build_hot_table(h);
        if (err) {  // leave the caller's arrays as they were handed in: undo the transpose
            if (launch_transpose(pl, false, h->stream) == cudaSuccess) cudaStreamSynchronize(h->stream);
            release_csr5_arrays(h);
            return err;
        }
    }
    h->convert_host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
#undef CUF
    h->format = CSR5B200_FORMAT_CSR5;
    return CSR5B200_SUCCESS;
}

int csr5b200_as_csr(csr5b200_handle_t h)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    if (h->format == CSR5B200_FORMAT_CSR) return CSR5B200_SUCCESS;
    if (h->format != CSR5B200_FORMAT_CSR5) return CSR5B200_UNKNOWN_FORMAT;
    if (h->pl.p > 0) {
        if (h->pl.hot_k > 0)
            CU(h, launch_hot_restore(h->pl.col, (long long)(h->pl.p - 1) * OMEGA * h->pl.sigma, h->pl.hot_col,
                                     h->tune.num_sms, h->stream));
        CU(h, launch_transpose(h->pl, false, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    release_csr5_arrays(h);
    h->format = CSR5B200_FORMAT_CSR;
    return CSR5B200_SUCCESS;
}

static int spmv_impl(csr5b200_handle_t h, double alpha, double beta, void *y, int n_dst, void *const *y_dst, int multicast)
{
    const ShardCtx *sh = nullptr;
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    if (h->format == CSR5B200_FORMAT_CSR) return CSR5B200_UNSUPPORTED_CSR_SPMV;
    if (h->format != CSR5B200_FORMAT_CSR5) return CSR5B200_UNKNOWN_FORMAT;
    if ((!y && h->pl.m > 0) || (!h->pl.x && h->pl.nnz > 0)) return CSR5B200_INVALID_ARGUMENT;
    if (n_dst < 0 || n_dst > CSR5B200_MAX_SCATTER || (n_dst > 0 && !y_dst)) return CSR5B200_INVALID_ARGUMENT;
    for (int k = 0; k < n_dst; k++)
        if (!y_dst[k]) return CSR5B200_INVALID_ARGUMENT;
    if (h->ignore_alpha) alpha = 1.0;
    if (n_dst > 0) {
        ShardCtx &s = h->shard;
        s.n_dst = n_dst;
        s.y_dst = y_dst;
        s.multicast = multicast;
        s.exchange = h->tune.exchange;
        sh = &s;
        if (beta != 0.0) return CSR5B200_INVALID_ARGUMENT;
        // fused scheme: carries are completed in y_local and re-sent from there, so the unicast destination list
        // must contain y_local; otherwise the rows would never reach it -- use the push scheme instead
        int exchange = s.exchange;
        if (exchange == 0)
            exchange = (!h->pl.needs_zero_fill && h->pl.m > 0 && (long long)h->pl.nnz / h->pl.m <= 64) ? 1 : 2;
        if (exchange == 1 && !multicast && h->pl.m > 0) {
            bool has_local = false;
            for (int k = 0; k < n_dst; k++) has_local |= y_dst[k] == y;
            if (!has_local) return CSR5B200_INVALID_ARGUMENT;
        }
    }
    cudaError_t e;
    h->tune.ev_begin = h->tune.ev_end = nullptr;
    // the main kernel is bracketed only when one is launched (p > 0, m > 0); the pair is kept only on success
    const bool timed = h->kernel_timing && h->pl.m > 0 && h->pl.p > 0 && h->ev_used + 2 <= 2 * MAX_TIMED_SPMV;
    if (timed) {
        while (h->ev.size() < h->ev_used + 2) {
            cudaEvent_t ev;
            CU(h, cudaEventCreate(&ev));
            h->ev.push_back(ev);
        }
        h->tune.ev_begin = h->ev[h->ev_used];
        h->tune.ev_end = h->ev[h->ev_used + 1];
    }
    h->launches_per_spmv = 0;
    const SpmvCall all;
    if (h->pl.value_bytes == 8)
        e = launch_spmv_part_f64(h->pl, h->tune, alpha, beta, static_cast<double *>(y), sh, all, h->stream,
                                 &h->kernel_in_use, &h->launches_per_spmv);
    else
        e = launch_spmv_part_f32(h->pl, h->tune, (float)alpha, (float)beta, static_cast<float *>(y), sh, all, h->stream,
                                 &h->kernel_in_use, &h->launches_per_spmv);
    h->tune.ev_begin = h->tune.ev_end = nullptr;
    if (e != cudaSuccess) return cuda_fail(h, e);
    if (timed) h->ev_used += 2;
    return CSR5B200_SUCCESS;
}

int csr5b200_spmv(csr5b200_handle_t h, double alpha, void *y) { return spmv_impl(h, alpha, 0.0, y, 0, nullptr, 0); }

int csr5b200_spmv_axpby(csr5b200_handle_t h, double alpha, double beta, void *y)
{
    return spmv_impl(h, alpha, beta, y, 0, nullptr, 0);
}

int csr5b200_spmv_scatter(csr5b200_handle_t h, double alpha, void *y_local, int n_dst, void *const *y_dst,
                          int dst_is_multicast)
{
    if (n_dst < 1 || (dst_is_multicast && n_dst != 1)) return CSR5B200_INVALID_ARGUMENT;
    return spmv_impl(h, alpha, 0.0, y_local, n_dst, y_dst, dst_is_multicast ? 1 : 0);
}

int csr5b200_destroy(csr5b200_handle_t h)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    int err = CSR5B200_SUCCESS;
    if (h->format == CSR5B200_FORMAT_CSR5) err = csr5b200_as_csr(h);
    cudaFree(h->x_stage);
    cudaFree(h->y_stage);
    h->x_stage = h->y_stage = nullptr;
    for (cudaEvent_t ev : h->ev) cudaEventDestroy(ev);
    h->ev.clear();
    h->ev_used = 0;
    for (int b = 0; b < 2; b++) {
        cudaFree(h->xb[b]); cudaFree(h->yb[b]);
        h->xb[b] = h->yb[b] = nullptr;
        if (h->e_in[b]) cudaEventDestroy(h->e_in[b]);
        if (h->e_comp[b]) cudaEventDestroy(h->e_comp[b]);
        if (h->e_out[b]) cudaEventDestroy(h->e_out[b]);
        h->e_in[b] = h->e_comp[b] = h->e_out[b] = nullptr;
    }
    h->shard = ShardCtx();
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    h->s_in = h->s_out = nullptr;
    h->batch_ready = false;
    release_exchange(h);
    free_pool(h);
    return err;
}

int csr5b200_free(csr5b200_handle_t h)
{
    if (!h) return CSR5B200_SUCCESS;
    const int err = csr5b200_destroy(h);
    delete h;
    return err;
}

int csr5b200_get_info(csr5b200_handle_t h, csr5b200_info *out)
{
    if (!h || !out) return CSR5B200_INVALID_ARGUMENT;
    const Plan &pl = h->pl;
    std::memset(out, 0, sizeof(*out));
    out->format = h->format;
    out->m = pl.m;
    out->n = pl.n;
    out->nnz = pl.nnz;
    out->value_bytes = pl.value_bytes;
    out->sigma = pl.sigma;
    out->bit_y_offset = pl.bit_y;
    out->bit_scansum_offset = pl.bit_ss;
    out->num_packet = pl.num_packet;
    out->p = pl.p;
    out->num_offsets = pl.num_offsets;
    out->tail_partition_start = pl.tail_start;
    out->needs_zero_fill = pl.needs_zero_fill;
    out->kernel_in_use = h->kernel_in_use;
    out->partition_pointer = pl.tile_ptr;
    out->partition_descriptor = pl.desc;
    out->partition_descriptor_offset_pointer = pl.desc_off_ptr;
    out->partition_descriptor_offset = pl.desc_off;
    out->calibrator = pl.calibrator;
    out->last_cuda_error = h->last_cuda_error;
    out->launches_per_spmv = h->launches_per_spmv;
    out->hot_columns = pl.hot_k;
    out->hot_coverage = pl.hot_coverage;
    for (int k = 0; k < 8; k++) out->convert_phase_ms[k] = h->convert_ms[k];
    out->convert_host_ms = h->convert_host_ms;
    out->convert_alloc_ms = h->convert_alloc_ms;
    out->exchange_transport = h->ex.last_transport;
    out->exchange_chunks = h->ex.last_chunks;
    out->has_carries = pl.has_carries;
    return CSR5B200_SUCCESS;
}

int csr5b200_get_kernel_times(csr5b200_handle_t h, float *ms, int capacity, int *count)
{
    if (!h || !count || (capacity > 0 && !ms)) return CSR5B200_INVALID_ARGUMENT;
    *count = 0;
    CU(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i + 1 < h->ev_used && *count < capacity; i += 2) {
        CU(h, cudaEventElapsedTime(ms + *count, h->ev[i], h->ev[i + 1]));
        ++*count;
    }
    h->ev_used = 0;
    return CSR5B200_SUCCESS;
}

int csr5b200_copy_meta_to_host(csr5b200_handle_t h, uint32_t *partition_pointer, uint32_t *partition_descriptor,
                               int32_t *partition_descriptor_offset_pointer, int32_t *partition_descriptor_offset,
                               void *calibrator)
{
    if (!h) return CSR5B200_INVALID_ARGUMENT;
    if (h->format != CSR5B200_FORMAT_CSR5) return CSR5B200_UNKNOWN_FORMAT;
    const Plan &pl = h->pl;
    CU(h, cudaStreamSynchronize(h->stream));
    if (pl.p == 0) return CSR5B200_SUCCESS;
    const cudaMemcpyKind k = cudaMemcpyDeviceToHost;
    if (partition_pointer) CU(h, cudaMemcpy(partition_pointer, pl.tile_ptr, (size_t)(pl.p + 1) * 4, k));
    if (partition_descriptor)
        CU(h, cudaMemcpy(partition_descriptor, pl.desc, (size_t)pl.p * OMEGA * pl.num_packet * 4, k));
    if (partition_descriptor_offset_pointer)
        CU(h, cudaMemcpy(partition_descriptor_offset_pointer, pl.desc_off_ptr, (size_t)(pl.p + 1) * 4, k));
    if (partition_descriptor_offset && pl.num_offsets > 0)
        CU(h, cudaMemcpy(partition_descriptor_offset, pl.desc_off, (size_t)pl.num_offsets * 4, k));
    if (calibrator) CU(h, cudaMemcpy(calibrator, pl.calibrator, (size_t)pl.p * pl.value_bytes, k));
    return CSR5B200_SUCCESS;
}

int csr5b200_spmv_host(csr5b200_handle_t h, double alpha, const void *x_host, void *y_host)
{
    if (!h || !x_host || !y_host) return CSR5B200_INVALID_ARGUMENT;
    if (h->format != CSR5B200_FORMAT_CSR5) return CSR5B200_UNSUPPORTED_CSR_SPMV;
    const size_t vb = (size_t)h->pl.value_bytes;
    if (!h->x_stage) CU(h, cudaMalloc(&h->x_stage, (size_t)(h->pl.n > 0 ? h->pl.n : 1) * vb));
    if (!h->y_stage) CU(h, cudaMalloc(&h->y_stage, (size_t)(h->pl.m > 0 ? h->pl.m : 1) * vb));
    CU(h, cudaMemcpyAsync(h->x_stage, x_host, (size_t)h->pl.n * vb, cudaMemcpyHostToDevice, h->stream));
    const void *saved_x = h->pl.x;
    h->pl.x = h->x_stage;
    const int err = csr5b200_spmv(h, alpha, h->y_stage);
    h->pl.x = saved_x;
    if (err) return err;
    CU(h, cudaMemcpyAsync(y_host, h->y_stage, (size_t)h->pl.m * vb, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return CSR5B200_SUCCESS;
}

int csr5b200_spmv_host_batch(csr5b200_handle_t h, double alpha, int count, const void *const *x_hosts,
                             void *const *y_hosts)
{
    if (!h || count < 0 || (count > 0 && (!x_hosts || !y_hosts))) return CSR5B200_INVALID_ARGUMENT;
    if (h->format != CSR5B200_FORMAT_CSR5) return CSR5B200_UNSUPPORTED_CSR_SPMV;
    if (count == 0) return CSR5B200_SUCCESS;
    const size_t vb = (size_t)h->pl.value_bytes;
    const size_t xbytes = (size_t)h->pl.n * vb, ybytes = (size_t)h->pl.m * vb;
    if (!h->batch_ready) {   // set only once every resource exists: a partial failure is retried from where it stopped
        if (!h->s_in) CU(h, cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
        if (!h->s_out) CU(h, cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            if (!h->xb[b]) CU(h, cudaMalloc(&h->xb[b], xbytes ? xbytes : 1));
            if (!h->yb[b]) CU(h, cudaMalloc(&h->yb[b], ybytes ? ybytes : 1));
            if (!h->e_in[b]) CU(h, cudaEventCreateWithFlags(&h->e_in[b], cudaEventDisableTiming));
            if (!h->e_comp[b]) CU(h, cudaEventCreateWithFlags(&h->e_comp[b], cudaEventDisableTiming));
            if (!h->e_out[b]) CU(h, cudaEventCreateWithFlags(&h->e_out[b], cudaEventDisableTiming));
        }
        h->batch_ready = true;
    }
    // work already queued on the handle's stream (e.g. a previous spmv) precedes the pipeline
    CU(h, cudaEventRecord(h->e_comp[0], h->stream));
    CU(h, cudaStreamWaitEvent(h->s_in, h->e_comp[0], 0));
    const void *saved_x = h->pl.x;
    int err = CSR5B200_SUCCESS;
    cudaError_t ce = cudaSuccess;
    auto ok = [&](cudaError_t e) { ce = e; return e == cudaSuccess; };
    for (int k = 0; k < count && !err && ce == cudaSuccess; k++) {
        const int b = k & 1;
        if (!x_hosts[k] || !y_hosts[k]) { err = CSR5B200_INVALID_ARGUMENT; break; }
        // upload x_k once the SpMV that last read this x buffer (k - 2) is done
        if (k >= 2 && !ok(cudaStreamWaitEvent(h->s_in, h->e_comp[b], 0))) break;
        if (!ok(cudaMemcpyAsync(h->xb[b], x_hosts[k], xbytes, cudaMemcpyHostToDevice, h->s_in))) break;
        if (!ok(cudaEventRecord(h->e_in[b], h->s_in))) break;
        // SpMV k after its upload and after the download that last read this y buffer (k - 2)
        if (!ok(cudaStreamWaitEvent(h->stream, h->e_in[b], 0))) break;
        if (k >= 2 && !ok(cudaStreamWaitEvent(h->stream, h->e_out[b], 0))) break;
        h->pl.x = h->xb[b];
        err = csr5b200_spmv(h, alpha, h->yb[b]);
        if (err) break;
        if (!ok(cudaEventRecord(h->e_comp[b], h->stream))) break;
        // download y_k
        if (!ok(cudaStreamWaitEvent(h->s_out, h->e_comp[b], 0))) break;
        if (!ok(cudaMemcpyAsync(y_hosts[k], h->yb[b], ybytes, cudaMemcpyDeviceToHost, h->s_out))) break;
        if (!ok(cudaEventRecord(h->e_out[b], h->s_out))) break;
    }
    h->pl.x = saved_x;
    cudaError_t e1 = cudaStreamSynchronize(h->s_in), e2 = cudaStreamSynchronize(h->stream),
                e3 = cudaStreamSynchronize(h->s_out);
    if (err) return err;
    if (ce != cudaSuccess) return cuda_fail(h, ce);
    if (e1 != cudaSuccess) return cuda_fail(h, e1);
    if (e2 != cudaSuccess) return cuda_fail(h, e2);
    if (e3 != cudaSuccess) return cuda_fail(h, e3);
    return CSR5B200_SUCCESS;
}

int csr5b200_call_anonymouslib(int m, int n, int nnz, const int *row_ptr_host, const int *col_host,
                               const void *val_host, const void *x_host, void *y_host, double alpha,
                               int value_bytes, int sigma)
{
    csr5b200_handle_t h = nullptr;
    int err = csr5b200_create(m, n, value_bytes, &h);
    if (err) return err;
    const size_t vb = (size_t)value_bytes;
    int *d_rp = nullptr, *d_col = nullptr;
    void *d_val = nullptr;
    auto cleanup = [&](int code) {
        csr5b200_free(h);
        cudaFree(d_rp);
        cudaFree(d_col);
        cudaFree(d_val);
        return code;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&d_rp, (size_t)(m + 1) * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&d_col, (size_t)(nnz > 0 ? nnz : 1) * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&d_val, (size_t)(nnz > 0 ? nnz : 1) * vb)) != cudaSuccess ||
        (e = cudaMemcpy(d_rp, row_ptr_host, (size_t)(m + 1) * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(d_col, col_host, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(d_val, val_host, (size_t)nnz * vb, cudaMemcpyHostToDevice)) != cudaSuccess) {
        h->last_cuda_error = (int)e;
        return cleanup(CSR5B200_CUDA_ERROR);
    }
    if ((err = csr5b200_input_csr(h, nnz, d_rp, d_col, d_val))) return cleanup(err);
    if ((err = csr5b200_set_sigma(h, sigma))) return cleanup(err);
    if ((err = csr5b200_warmup(h))) return cleanup(err);
    if ((err = csr5b200_as_csr5(h))) return cleanup(err);
    if ((err = csr5b200_spmv_host(h, alpha, x_host, y_host))) return cleanup(err);
    return cleanup(CSR5B200_SUCCESS);
}

const char *csr5b200_version(void) { return "csr5-b200 0.2 (sm_100a)"; }

const char *csr5b200_error_string(int code)
{
    switch (code) {
        case CSR5B200_SUCCESS: return "success";
        case CSR5B200_UNKNOWN_FORMAT: return "unknown format (inputCSR not called)";
        case CSR5B200_UNSUPPORTED_CSR5_OMEGA: return "unsupported CSR5 omega/sigma bit budget";
        case CSR5B200_CSR_TO_CSR5_FAILED: return "CSR -> CSR5 conversion failed (sigma outside [4, 32]?)";
        case CSR5B200_UNSUPPORTED_CSR_SPMV: return "spmv on CSR format: call asCSR5 first";
        case CSR5B200_UNSUPPORTED_VALUE_TYPE: return "unsupported value type (use 4 or 8 bytes)";
        case CSR5B200_CUDA_ERROR: return "CUDA runtime error (see csr5b200_info.last_cuda_error)";
        case CSR5B200_INVALID_ARGUMENT: return "invalid argument";
        case CSR5B200_EXCHANGE_TIMEOUT: return "a device-side barrier of the sharded step timed out waiting for a peer";
        default: return "unrecognised error code";
    }
}

}  // extern "C"
