// csr5_coo.cu -- COO -> CSR on the device with the semantics of the reference's loader (csr5b200_coo_to_csr).
//
// Replaces the host loops of CSR5_cuda/main.cu:211-306 (the same code is in every backend's main):
//   * symmetric / hermitian files: every off-diagonal entry (i, j) also yields (j, i), emitted right after it
//     (main.cu:239-246, 271-289);
//   * counting sort by row that keeps the emission order inside a row -- columns are NOT sorted, duplicates are
//     kept (main.cu:248-306: per-row counters, exclusive scan, sequential fill).
// The sequential fill is what fixes the order inside a row, so a parallel version has to be a STABLE sort by row:
// here a least-significant-digit radix sort (8 bits per pass, ceil(log2(m) / 8) passes) over (row, source index)
// pairs -- per-block digit histograms, one device-wide exclusive scan (digit-major), and an order-preserving
// scatter (warp match + per-warp running digit offsets) -- followed by one gather of col / val.  Everything is
// coalesced streaming except the final gather; no atomics decide an order.
#include <vector>

#include "csr5_internal.h"

namespace csr5 {
namespace {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int RS_WARPS = 8;
constexpr int RS_ITEMS = 16;                       // 32-element rounds per warp
constexpr int RS_TILE = RS_WARPS * 32 * RS_ITEMS;  // elements per block
constexpr uint32_t TRANSPOSED = 0x80000000u;       // source index tag: the (j, i) copy of entry (i, j)

// flag[i] = 1 if entry i is mirrored; also validates the indices (bad[0] != 0 on failure)
__global__ void __launch_bounds__(256)
coo_flag_kernel(const int *__restrict__ rows, const int *__restrict__ cols, int nnz, int m, int n, int symmetric,
                int *__restrict__ flag, int *__restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nnz) return;
    int f = 0;
    if (i < nnz) {
        const int r = rows[i], c = cols[i];
        if (r < 0 || r >= m || c < 0 || c >= n) *bad = 1;
        if (symmetric && r != c) {
            f = 1;
            if (c >= m || r >= n) *bad = 1;   // the mirrored entry must fit too
        }
    }
    flag[i] = f;
}

// expanded list E in emission order: position of entry i is i + (mirrored entries before i)
__global__ void __launch_bounds__(256)
coo_expand_kernel(const int *__restrict__ rows, const int *__restrict__ cols, int nnz, const int *__restrict__ before,
                  uint32_t *__restrict__ key, uint32_t *__restrict__ src, int *__restrict__ row_cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int e = i + before[i];
    const int r = rows[i], c = cols[i];
    key[e] = (uint32_t)r;
    src[e] = (uint32_t)i;
    atomicAdd(row_cnt + r, 1);
    if (before[i + 1] != before[i]) {
        key[e + 1] = (uint32_t)c;
        src[e + 1] = (uint32_t)i | TRANSPOSED;
        atomicAdd(row_cnt + c, 1);
    }
}

__global__ void __launch_bounds__(256)
rs_hist_kernel(const uint32_t *__restrict__ key, int n, int shift, int *__restrict__ hist, int nblocks)
{
    __shared__ int s[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) s[t] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE;
    for (int k = threadIdx.x; k < RS_TILE; k += blockDim.x) {
        const long long j = base + k;
        if (j < n) atomicAdd(&s[(key[j] >> shift) & 255u], 1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 256; t += blockDim.x) hist[(size_t)t * nblocks + blockIdx.x] = s[t];
}

// Order-preserving scatter of one radix pass.  Warp w owns the w-th run of 32 * RS_ITEMS consecutive elements of
// the block's tile and walks it 32 at a time; inside a round the rank among equal digits comes from
// __match_any_sync, across rounds from the warp's running per-digit offset, across warps and blocks from prefix sums.
__global__ void __launch_bounds__(RS_WARPS * 32)
rs_scatter_kernel(const uint32_t *__restrict__ key_in, const uint32_t *__restrict__ src_in, uint32_t *__restrict__ key_out,
                  uint32_t *__restrict__ src_out, int n, int shift, const int *__restrict__ offs, int nblocks)
{
    __shared__ int s_cnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < RS_WARPS * 256; t += blockDim.x) (&s_cnt[0][0])[t] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE + (long long)w * 32 * RS_ITEMS;
    const unsigned lt = (1u << lane) - 1u;
    // digit counts of this warp's run
    for (int it = 0; it < RS_ITEMS; it++) {
        const long long j = base + it * 32 + lane;
        const bool ok = j < n;
        const unsigned d = ok ? ((key_in[j] >> shift) & 255u) : 0xffffu;
        const unsigned peers = __match_any_sync(FULLM, d);
        if (ok && (peers & lt) == 0) s_cnt[w][d] += __popc(peers);   // lowest lane of each digit group
        __syncwarp();
    }
    __syncthreads();
    // digit d: global offset of the block, then prefix over the warps
    for (int d = threadIdx.x; d < 256; d += blockDim.x) {
        int run = offs[(size_t)d * nblocks + blockIdx.x];
        for (int ww = 0; ww < RS_WARPS; ww++) {
            const int c = s_cnt[ww][d];
            s_cnt[ww][d] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int it = 0; it < RS_ITEMS; it++) {
        const long long j = base + it * 32 + lane;
        const bool ok = j < n;
        uint32_t k = 0, s = 0;
        if (ok) { k = key_in[j]; s = src_in[j]; }
        const unsigned d = ok ? ((k >> shift) & 255u) : 0xffffu;
        const unsigned peers = __match_any_sync(FULLM, d);
        int pos = 0;
        if (ok) pos = s_cnt[w][d] + __popc(peers & lt);
        __syncwarp();
        if (ok) {
            key_out[pos] = k;
            src_out[pos] = s;
            if ((peers & lt) == 0) s_cnt[w][d] += __popc(peers);
        }
        __syncwarp();
    }
}

template <typename VT>
__global__ void __launch_bounds__(256)
coo_gather_kernel(const uint32_t *__restrict__ src, int nnz_out, const int *__restrict__ rows, const int *__restrict__ cols,
                  const VT *__restrict__ vals, int *__restrict__ col_out, VT *__restrict__ val_out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz_out) return;
    const uint32_t s = src[j];
    const int i = (int)(s & ~TRANSPOSED);
    col_out[j] = (s & TRANSPOSED) ? rows[i] : cols[i];
    val_out[j] = vals ? vals[i] : (VT)1;   // pattern files: every entry is 1.0 (main.cu:232-235)
}

}  // namespace
}  // namespace csr5

using namespace csr5;

extern "C" {

int csr5b200_coo_to_csr(int m, int n, int nnz, const int *rows, const int *cols, const void *vals, int value_bytes,
                        int symmetric, int *row_ptr, int *col_out, void *val_out, int capacity, int *nnz_out,
                        void *cuda_stream)
{
    if (m < 0 || n < 0 || nnz < 0 || !row_ptr || !nnz_out || (nnz > 0 && (!rows || !cols))) return CSR5B200_INVALID_ARGUMENT;
    if (value_bytes != 4 && value_bytes != 8) return CSR5B200_UNSUPPORTED_VALUE_TYPE;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int *before = nullptr, *bad = nullptr, *hist = nullptr;
    uint32_t *key[2] = {nullptr, nullptr}, *src[2] = {nullptr, nullptr};
    void *scratch = nullptr;
    int code = CSR5B200_SUCCESS;
    auto done = [&](int c) {
        cudaFree(before); cudaFree(bad); cudaFree(hist); cudaFree(scratch);
        cudaFree(key[0]); cudaFree(key[1]); cudaFree(src[0]); cudaFree(src[1]);
        return c;
    };
#define CUC(call)                                                         \
    do {                                                                  \
        if ((call) != cudaSuccess) return done(CSR5B200_CUDA_ERROR);      \
    } while (0)
    CUC(cudaMemsetAsync(row_ptr, 0, (size_t)(m + 1) * sizeof(int), st));
    if (nnz == 0) {
        CUC(cudaStreamSynchronize(st));
        *nnz_out = 0;
        return done(CSR5B200_SUCCESS);
    }
    // 1. which entries are mirrored, and where every entry lands in the emission order
    CUC(cudaMalloc(&before, (size_t)(nnz + 1) * sizeof(int)));
    CUC(cudaMalloc(&bad, sizeof(int)));
    CUC(cudaMemsetAsync(bad, 0, sizeof(int), st));
    coo_flag_kernel<<<(nnz + 1 + 255) / 256, 256, 0, st>>>(rows, cols, nnz, m, n, symmetric ? 1 : 0, before, bad);
    {
        const size_t sb = scan_scratch_bytes(nnz + 1);
        CUC(cudaMalloc(&scratch, sb));
        CUC(launch_exclusive_scan(before, nnz + 1, scratch, sb, st));
    }
    int mirrored = 0, is_bad = 0;
    CUC(cudaMemcpyAsync(&mirrored, before + nnz, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUC(cudaMemcpyAsync(&is_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUC(cudaStreamSynchronize(st));
    if (is_bad) return done(CSR5B200_INVALID_ARGUMENT);
    const long long total = (long long)nnz + mirrored;
    *nnz_out = (int)total;
    if (total > 0x7fffffffLL) return done(CSR5B200_INVALID_ARGUMENT);
    if (!col_out || !val_out || capacity < total) return done(capacity < total && col_out ? CSR5B200_INVALID_ARGUMENT : CSR5B200_SUCCESS);
    const int N = (int)total;
    // 2. (row, source) pairs in emission order + row counts
    for (int b = 0; b < 2; b++) {
        CUC(cudaMalloc(&key[b], (size_t)N * sizeof(uint32_t)));
        CUC(cudaMalloc(&src[b], (size_t)N * sizeof(uint32_t)));
    }
    coo_expand_kernel<<<(nnz + 255) / 256, 256, 0, st>>>(rows, cols, nnz, before, key[0], src[0], row_ptr);
    CUC(cudaGetLastError());
    {
        cudaFree(scratch);
        scratch = nullptr;
        const size_t sb = scan_scratch_bytes(m + 1);
        CUC(cudaMalloc(&scratch, sb));
        CUC(launch_exclusive_scan(row_ptr, m + 1, scratch, sb, st));   // counts -> row_ptr (main.cu:248-258)
    }
    // 3. stable LSD radix sort by row
    int bits = 1;
    while ((1LL << bits) < m) bits++;
    const int nblocks = (N + RS_TILE - 1) / RS_TILE;
    CUC(cudaMalloc(&hist, (size_t)256 * nblocks * sizeof(int) + sizeof(int)));
    {
        cudaFree(scratch);
        scratch = nullptr;
        const size_t sb = scan_scratch_bytes(256 * nblocks);
        CUC(cudaMalloc(&scratch, sb));
        int cur = 0;
        for (int shift = 0; shift < bits; shift += 8) {
            rs_hist_kernel<<<nblocks, 256, 0, st>>>(key[cur], N, shift, hist, nblocks);
            CUC(launch_exclusive_scan(hist, 256 * nblocks, scratch, sb, st));
            rs_scatter_kernel<<<nblocks, RS_WARPS * 32, 0, st>>>(key[cur], src[cur], key[cur ^ 1], src[cur ^ 1], N, shift, hist,
                                                                 nblocks);
            CUC(cudaGetLastError());
            cur ^= 1;
        }
        // 4. col / val in CSR order
        if (value_bytes == 8)
            coo_gather_kernel<double><<<(N + 255) / 256, 256, 0, st>>>(src[cur], N, rows, cols, static_cast<const double *>(vals),
                                                                     col_out, static_cast<double *>(val_out));
        else
            coo_gather_kernel<float><<<(N + 255) / 256, 256, 0, st>>>(src[cur], N, rows, cols, static_cast<const float *>(vals),
                                                                    col_out, static_cast<float *>(val_out));
        CUC(cudaGetLastError());
    }
    CUC(cudaStreamSynchronize(st));
#undef CUC
    return done(code);
}

}  // extern "C"
