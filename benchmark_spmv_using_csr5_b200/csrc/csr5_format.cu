// csr5_format.cu -- CSR -> CSR5 conversion (and back) for sm_100a.
//
// Produces word-for-word the arrays the reference's CSR5_cuda handle holds after asCSR5()
// (tile_ptr, tile_desc, desc_offset_ptr, desc_offset, transposed col/val), for
// every tile the reference's SpMV ever reads (t < p - 1), but with a different decomposition:
//
//   reference (format_cuda.h)                         here
//   s1 binary search per tile boundary  (21-42)       tile_ptr_kernel             (same search)
//   s2 block per tile, empty-row probe  (44-95)   \
//   desc s1: one global atomicOr per row (129-159) |   tile_desc_kernel: one warp per tile walks the
//   desc s2: sigma-step bit loop per lane (161-267)/   tile's slice of row_ptr ONCE (coalesced), builds
//                                                      the flags with shared-memory atomics, detects
//                                                      empty rows, and derives y_offset / seg_offset
//                                                      with popc / ballot / ffs -- no global atomics,
//                                                      no descriptor memset, no second pass
//   s3 single 256-thread block scan (269-300)         3-phase device-wide exclusive scan
//   offset kernel: binary search per flag (362-499)   desc_offset_kernel: second coalesced walk of the
//                                                      tile's rows; each non-empty row start knows its
//                                                      own slot from the packed flags (popc)
//   transpose: 2 launches, sigma-templated (525-744)  one launch moves col and val together
#include "csr5_internal.h"

namespace csr5 {

namespace {

constexpr unsigned FULL = 0xffffffffu;

// number of entries of the sorted array a[0..n) that are <= key  (utils_cuda.h:25-53)
__device__ __forceinline__ int count_le(const int *__restrict__ a, int n, int key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// first index i in [0, n) with a[i] >= key (n if none)
__device__ __forceinline__ int first_ge(const int *__restrict__ a, int n, long long key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((long long)__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// A tile whose rows [start, stop] number more than this is not walked row by row: long runs of empty rows (R-MAT:
// half of all rows) would serialise one warp over up to millions of rows.  Such a tile looks its <= omega * sigma
// positions up by binary search instead (the reference's own formulation, format_cuda.h:362-422), which is bounded
// by sigma * log2(span) probes per lane whatever the span.
constexpr uint32_t WALK_LIMIT = 2048;

// tile_ptr[t] = row that holds nnz index min(t * omega * sigma, nnz), last such row on ties.
__global__ void tile_ptr_kernel(const int *__restrict__ row_ptr, uint32_t *__restrict__ tile_ptr,
                                int *__restrict__ desc_off_ptr, int sigma, int p, int m, int nnz)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > p) return;
    long long b = (long long)t * sigma * OMEGA;
    if (b > nnz) b = nnz;
    tile_ptr[t] = (uint32_t)(count_le(row_ptr, m + 1, (int)b) - 1);
    desc_off_ptr[t] = 0;
}

template <int WPB>
__global__ void __launch_bounds__(WPB * 32)
tile_desc_kernel(const int *__restrict__ row_ptr, uint32_t *tile_ptr, uint32_t *__restrict__ desc,
                 int *__restrict__ desc_off_ptr, int *__restrict__ dev_flags, int sigma, int p, int m,
                 int bit_y, int bit_all, int num_packet)
{
    __shared__ uint32_t s_flags[WPB][OMEGA];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int t = blockIdx.x * WPB + w;
    if (t >= p) return;  // warp-uniform; no block-wide barrier below

    // Neighbouring warps may be OR-ing bit 31 into these words right now; both values mask alike.
    const uint32_t start = *(volatile uint32_t *)(tile_ptr + t) & ROW_MASK;
    const uint32_t stop = *(volatile uint32_t *)(tile_ptr + t + 1) & ROW_MASK;
    const long long base = (long long)t * sigma * OMEGA;
    const int tile = sigma * OMEGA;
    uint32_t *td = desc + (size_t)t * OMEGA * num_packet;

    if (start == stop) {
        // One row covers the whole tile (fast track).  The only row start that can fall inside is
        // one exactly on the tile boundary; the reference keeps that raw flag (format_cuda.h:187).
        uint32_t w0 = 0;
        if (lane == 0) {
            if ((long long)row_ptr[start] == base) w0 = 1u << (31 - bit_all);
            else if (t > 0) dev_flags[1] = 1;   // the tile continues a row: some SpMV carry exists
        }
        td[lane] = w0;
        if (num_packet > 1) td[OMEGA + lane] = 0;
        return;
    }

    bool dirty = false;
    uint32_t f = 0;  // bit i = element (lane, i) starts a row
    if (stop - start <= WALK_LIMIT) {
        s_flags[w][lane] = 0;
        __syncwarp();
        for (uint32_t r0 = start; r0 <= stop; r0 += 32) {
            const uint32_t r = r0 + lane;
            if (r <= stop && r < (uint32_t)m) {
                const int o = row_ptr[r];
                const int o1 = row_ptr[r + 1];
                if (r < stop && o == o1) dirty = true;  // rows [start, stop) as format_cuda.h:72-84
                const long long pos = (long long)o - base;
                if (pos >= 0 && pos < tile) {
                    const int ps = (int)pos;
                    atomicOr(&s_flags[w][ps / sigma], 1u << (ps % sigma));
                }
            }
        }
        __syncwarp();
        f = s_flags[w][lane];
    } else {
        // rows r in [start, R], R = min(stop, m - 1); row_ptr[R + 1] exists.  Position `pos` carries a flag iff some
        // row has row_ptr == base + pos; a second row with the same value means the first one is empty, and it lies
        // below `stop` (the last row of the run is <= R), i.e. the tile is dirty.
        const uint32_t R = stop < (uint32_t)m ? stop : (uint32_t)m - 1;
        const int *a = row_ptr + start;
        const int cnt = (int)(R - start) + 2;   // a[0 .. cnt): rows start .. R + 1
        for (int i = 0; i <= sigma; i++) {
            // lane l looks up its own sigma positions; position `tile` (the next tile's first element: empty rows
            // that share it are still rows of this tile) is lane 31's extra round
            if (i == sigma && lane != 31) break;
            const long long v = base + (long long)lane * sigma + i;
            const int k = first_ge(a, cnt, v);
            if (k < cnt - 1 && (long long)a[k] == v) {
                if (i < sigma) f |= 1u << i;
                if ((long long)a[k + 1] == v) dirty = true;
            }
        }
    }
    dirty = __any_sync(FULL, dirty);

    int y_off = 0, seg_off = 0, total = 0;
    if (t < p - 1) {
        const uint32_t ff = f | (lane == 0 ? 1u : 0u);  // lane 0 always opens a segment
        const int segn = __popc(ff);
        int incl = segn;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int nb = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += nb;
        }
        total = __shfl_sync(FULL, incl, 31);
        y_off = lane ? incl - segn - 1 : 0;
        const uint32_t present = __ballot_sync(FULL, ff != 0);
        if (ff) {
            const uint32_t following = lane == 31 ? 0u : present >> (lane + 1);
            seg_off = following ? __ffs(following) - 1 : 31 - lane;
        }
    }
    // [ y_offset : bit_y ][ seg_offset : bit_ss ][ flag 0 .. flag sigma-1 ], MSB first, over
    // num_packet 32-bit words
    const unsigned long long word = ((unsigned long long)(uint32_t)y_off << (64 - bit_y)) |
                                    ((unsigned long long)(uint32_t)seg_off << (64 - bit_all)) |
                                    ((unsigned long long)__brev(f) << (32 - bit_all));
    td[lane] = (uint32_t)(word >> 32);
    if (num_packet > 1) td[OMEGA + lane] = (uint32_t)word;

    if (lane == 0 && t > 0 && !(f & 1u)) dev_flags[1] = 1;   // first element is not a row start (tail tile included)
    if (lane == 0 && dirty) {
        tile_ptr[t] = start | MSB;
        if (t < p - 1) {
            desc_off_ptr[t] = total;
            dev_flags[0] = 1;
        }
    }
}

// ---- device-wide exclusive scan of (p + 1) ints, in place ------------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *s_warp, int *block_total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int nb = __shfl_up_sync(FULL, incl, d);
        if (lane >= d) incl += nb;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        int ws = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int nb = __shfl_up_sync(FULL, ws, d);
            if (lane >= d) ws += nb;
        }
        s_warp[lane] = ws;  // inclusive over warps
    }
    __syncthreads();
    const int warp_excl = w ? s_warp[w - 1] : 0;
    *block_total = s_warp[31];
    return warp_excl + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_reduce_kernel(const int *__restrict__ data, int n, int *__restrict__ block_sums)
{
    __shared__ int s_warp[32];
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) v += data[base + k];
    int total;
    block_exclusive_scan(v, s_warp, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(int *block_sums, int nb)
{
    __shared__ int s_warp[32];
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += SCAN_THREADS) {
        const int i = b0 + threadIdx.x;
        const int v = i < nb ? block_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan(v, s_warp, &total);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(int *__restrict__ data, int n, const int *__restrict__ block_sums)
{
    __shared__ int s_warp[32];
    const int base = blockIdx.x * SCAN_CHUNK + threadIdx.x * SCAN_ITEMS;
    int item[SCAN_ITEMS];
    int v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        item[k] = base + k < n ? data[base + k] : 0;
        v += item[k];
    }
    int total;
    int run = block_exclusive_scan(v, s_warp, &total) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) data[base + k] = run;
        run += item[k];
    }
}

// Empty-row table of dirty tiles: slot of a real flag at (lane, i) is
// desc_off_ptr[t] + y_offset[lane] + (#flags of that lane before i, not counting lane 0's bit 0);
// the value is the starting row's index relative to row_start + 1.
template <int WPB>
__global__ void __launch_bounds__(WPB * 32)
desc_offset_kernel(const int *__restrict__ row_ptr, const uint32_t *__restrict__ tile_ptr,
                   const uint32_t *__restrict__ desc, const int *__restrict__ desc_off_ptr,
                   int *__restrict__ desc_off, int sigma, int p, int bit_y, int bit_all, int num_packet)
{
    __shared__ uint32_t s_flags[WPB][OMEGA];
    __shared__ int s_yoff[WPB][OMEGA];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int t = blockIdx.x * WPB + w;
    if (t >= p - 1) return;
    const uint32_t raw = tile_ptr[t];
    if (!(raw & MSB)) return;
    const uint32_t start = raw & ROW_MASK;
    const uint32_t stop = tile_ptr[t + 1] & ROW_MASK;
    const long long base = (long long)t * sigma * OMEGA;
    const int tile = sigma * OMEGA;
    const uint32_t *td = desc + (size_t)t * OMEGA * num_packet;

    const uint32_t w0 = td[lane];
    const uint32_t w1 = num_packet > 1 ? td[OMEGA + lane] : 0u;
    const unsigned long long word = ((unsigned long long)w0 << 32) | w1;
    uint32_t f = __brev((uint32_t)(word >> (32 - bit_all)));
    if (sigma < 32) f &= (1u << sigma) - 1u;
    s_flags[w][lane] = f;
    s_yoff[w][lane] = (int)(w0 >> (32 - bit_y));
    __syncwarp();

    const int ob = desc_off_ptr[t];
    if (stop - start > WALK_LIMIT) {
        // bounded form for tiles that span long runs of empty rows: every flagged element finds the row that starts
        // there -- the LAST row with that row_ptr value -- by binary search (format_cuda.h:362-422)
        const int *a = row_ptr + start + 1;
        const int cnt = (int)(stop - start);     // rows start + 1 .. stop
        uint32_t mine = f;
        if (lane == 0) mine &= ~1u;
        const int yo = (int)(w0 >> (32 - bit_y));
        int k = 0;
        while (mine) {
            const int i = __ffs(mine) - 1;
            mine &= mine - 1;
            const long long v = base + (long long)lane * sigma + i;
            const int r = first_ge(a, cnt, v + 1) - 1;   // last row (relative to start + 1) with row_ptr <= v
            desc_off[ob + yo + k] = r;
            k++;
        }
        return;
    }
    for (uint32_t r0 = start + 1; r0 <= stop; r0 += 32) {
        const uint32_t r = r0 + lane;
        if (r <= stop) {
            const int o = row_ptr[r];
            const long long pos = (long long)o - base;
            if (pos >= 0 && pos < tile && row_ptr[r + 1] > o) {  // the non-empty row that starts here
                const int ps = (int)pos;
                const int l = ps / sigma, i = ps - l * sigma;
                uint32_t before = s_flags[w][l] & ((1u << i) - 1u);
                if (l == 0) before &= ~1u;
                if (l || i) desc_off[ob + s_yoff[w][l] + __popc(before)] = (int)(r - start - 1);
            }
        }
    }
}

// In-place transpose of one omega x sigma tile of col and val per CTA.  CSR order keeps a lane's
// sigma elements contiguous (lane * sigma + i); CSR5 order is i * 32 + lane.  A tile is skipped
// when the RAW words tile_ptr[t] == tile_ptr[t+1] (format_cuda.h:540).
//
// Replaces aosoa_transpose_kernel_smem (format_cuda.h:525-585: two launches, 29 sigma instantiations): one launch moves
// col and val together, sigma is a run-time value.  Every warp access is one fully coalesced 128- / 256-byte row and
// every shared-memory access is conflict-free (odd row stride).  Measured on B200 (profiles/r02_transpose_variants.txt):
// C2 0.61 ms = 6.3 TB/s (0.96 of the copy peak), C4 2.59 ms = 5.4 TB/s (0.83).  Two re-designs were measured and
// dropped: 16-byte vector global accesses with a multiply-high instead of the division (C4 2.66 ms: the scalar
// shared-memory side becomes 2-way conflicted), and the same with four tiles per 256-thread CTA (C4 2.93 ms,
// C2 0.70 ms).
template <typename VT>
__global__ void __launch_bounds__(128)
transpose_kernel(int *__restrict__ col, VT *__restrict__ val, const uint32_t *__restrict__ tile_ptr,
                 int sigma, bool r2c)
{
    const unsigned t = blockIdx.x;
    if (tile_ptr[t] == tile_ptr[t + 1]) return;
    extern __shared__ __align__(16) unsigned char smem[];
    const int stride = sigma | 1;  // odd stride: conflict-free column reads
    VT *sv = reinterpret_cast<VT *>(smem);
    int *sc = reinterpret_cast<int *>(sv + OMEGA * stride);
    const int tile = OMEGA * sigma;
    const size_t base = (size_t)t * tile;

    for (int idx = threadIdx.x; idx < tile; idx += 128) {
        int l, i;
        if (r2c) { l = idx / sigma; i = idx - l * sigma; } else { i = idx >> 5; l = idx & 31; }
        sv[l * stride + i] = val[base + idx];
        sc[l * stride + i] = col[base + idx];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < tile; idx += 128) {
        int l, i;
        if (r2c) { i = idx >> 5; l = idx & 31; } else { l = idx / sigma; i = idx - l * sigma; }
        val[base + idx] = sv[l * stride + i];
        col[base + idx] = sc[l * stride + i];
    }
}

// ---- hot-column table -------------------------------------------------------------------------------
// cnt[c] = number of references to column c among the first `limit` non-zeros (the CSR5 tiles).
__global__ void __launch_bounds__(256) hot_count_kernel(const int *__restrict__ col, long long limit, int *cnt)
{
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < limit; k += (long long)gridDim.x * blockDim.x)
        atomicAdd(cnt + col[k], 1);
}

// out2[0] = #columns with cnt >= threshold, out2[1] = references they receive
__global__ void __launch_bounds__(256)
hot_count_ge_kernel(const int *__restrict__ cnt, int n, int threshold, unsigned long long *out2)
{
    unsigned long long cols = 0, refs = 0;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int v = cnt[c];
        if (v >= threshold) { cols++; refs += (unsigned long long)v; }
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) {
        cols += __shfl_xor_sync(FULL, cols, w);
        refs += __shfl_xor_sync(FULL, refs, w);
    }
    if ((threadIdx.x & 31) == 0 && cols) { atomicAdd(out2, cols); atomicAdd(out2 + 1, refs); }
}

// One pass over the column reference counts: cols[v] = number of columns with min(cnt, HOT_BINS - 1) == v and
// refs[v] = the references they receive.  The host picks the threshold from the two histograms (one read-back
// instead of a bisection of ~30 synchronising launches).  Low counts -- almost every column of a power-law matrix
// -- are accumulated in shared memory first.
constexpr int HOT_BINS = 65536;
constexpr int HOT_SMEM_BINS = 2048;

__global__ void __launch_bounds__(256)
hot_hist_kernel(const int *__restrict__ cnt, int n, unsigned int *__restrict__ cols, unsigned long long *__restrict__ refs)
{
    __shared__ unsigned int s_cols[HOT_SMEM_BINS];
    for (int t = threadIdx.x; t < HOT_SMEM_BINS; t += blockDim.x) s_cols[t] = 0;
    __syncthreads();
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (long long)gridDim.x * blockDim.x) {
        const int v = cnt[c];
        if (v <= 0) continue;
        if (v < HOT_SMEM_BINS) {
            atomicAdd(&s_cols[v], 1u);
        } else {
            const int b = v < HOT_BINS ? v : HOT_BINS - 1;
            atomicAdd(cols + b, 1u);
            atomicAdd(refs + b, (unsigned long long)v);
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < HOT_SMEM_BINS; t += blockDim.x) {
        const unsigned int k = s_cols[t];
        if (k) {
            atomicAdd(cols + t, k);
            atomicAdd(refs + t, (unsigned long long)k * (unsigned long long)t);   // every column of bin t has count t
        }
    }
}

__global__ void __launch_bounds__(256) hot_flags_kernel(const int *__restrict__ cnt, int n, int threshold, int *slot)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= n) slot[c] = (c < n && cnt[c] >= threshold) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
hot_fill_kernel(const int *__restrict__ cnt, int n, int threshold, const int *__restrict__ slot, int *__restrict__ hot_col)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && cnt[c] >= threshold) hot_col[slot[c]] = c;
}

// A hot column index is replaced by (bit 31 | slot); asCSR() undoes it from hot_col[].
__global__ void __launch_bounds__(256)
hot_rewrite_kernel(int *col, long long limit, const int *__restrict__ cnt, int threshold, const int *__restrict__ slot)
{
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < limit; k += (long long)gridDim.x * blockDim.x) {
        const int c = col[k];
        if (__ldg(cnt + c) >= threshold) col[k] = (int)(MSB | (uint32_t)__ldg(slot + c));
    }
}

__global__ void __launch_bounds__(256) hot_restore_kernel(int *col, long long limit, const int *__restrict__ hot_col)
{
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < limit; k += (long long)gridDim.x * blockDim.x) {
        const int c = col[k];
        if (c < 0) col[k] = __ldg(hot_col + (c & (int)ROW_MASK));
    }
}

__global__ void warmup_kernel(int *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0 && out) *out = 0;
}

}  // namespace

cudaError_t launch_tile_ptr(const Plan &pl, cudaStream_t stream)
{
    const int threads = 256;
    const int blocks = (pl.p + 1 + threads - 1) / threads;
    tile_ptr_kernel<<<blocks, threads, 0, stream>>>(pl.row_ptr, pl.tile_ptr, pl.desc_off_ptr, pl.sigma,
                                                    pl.p, pl.m, pl.nnz);
    return cudaGetLastError();
}

cudaError_t launch_tile_desc(const Plan &pl, cudaStream_t stream)
{
    constexpr int WPB = 8;
    const int blocks = (pl.p + WPB - 1) / WPB;
    tile_desc_kernel<WPB><<<blocks, WPB * 32, 0, stream>>>(pl.row_ptr, pl.tile_ptr, pl.desc,
                                                           pl.desc_off_ptr, pl.dev_flags, pl.sigma, pl.p,
                                                           pl.m, pl.bit_y, pl.bit_y + pl.bit_ss,
                                                           pl.num_packet);
    return cudaGetLastError();
}

size_t scan_scratch_bytes(int p)
{
    const int nb = (p + 1 + SCAN_CHUNK - 1) / SCAN_CHUNK;
    return (size_t)nb * sizeof(int);
}

cudaError_t launch_exclusive_scan(int *data, int n, void *scratch, size_t scratch_bytes, cudaStream_t stream)
{
    const int nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (scratch_bytes < (size_t)nb * sizeof(int)) return cudaErrorInvalidValue;
    int *block_sums = static_cast<int *>(scratch);
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, stream>>>(data, n, block_sums);
    scan_block_sums_kernel<<<1, SCAN_THREADS, 0, stream>>>(block_sums, nb);
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, stream>>>(data, n, block_sums);
    return cudaGetLastError();
}

cudaError_t launch_scan_offsets(const Plan &pl, void *scratch, size_t scratch_bytes, cudaStream_t stream)
{
    return launch_exclusive_scan(pl.desc_off_ptr, pl.p + 1, scratch, scratch_bytes, stream);
}

// ---- hot-column table (no reference counterpart) ---------------------------------
cudaError_t launch_hot_count(const int *col, long long limit, int *cnt, int num_sms, cudaStream_t stream)
{
    hot_count_kernel<<<num_sms * 16, 256, 0, stream>>>(col, limit, cnt);
    return cudaGetLastError();
}

cudaError_t launch_hot_count_ge(const int *cnt, int n, int threshold, unsigned long long *out2, int num_sms,
                                cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    hot_count_ge_kernel<<<num_sms * 4, 256, 0, stream>>>(cnt, n, threshold, out2);
    return cudaGetLastError();
}

int hot_hist_bins() { return HOT_BINS; }

cudaError_t launch_hot_hist(const int *cnt, int n, unsigned int *cols, unsigned long long *refs, int num_sms,
                            cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(cols, 0, HOT_BINS * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(refs, 0, HOT_BINS * sizeof(unsigned long long), stream)) != cudaSuccess) return e;
    hot_hist_kernel<<<num_sms * 4, 256, 0, stream>>>(cnt, n, cols, refs);
    return cudaGetLastError();
}

cudaError_t launch_hot_flags(const int *cnt, int n, int threshold, int *slot, cudaStream_t stream)
{
    hot_flags_kernel<<<(n + 1 + 255) / 256, 256, 0, stream>>>(cnt, n, threshold, slot);
    return cudaGetLastError();
}

cudaError_t launch_hot_assign(const int *cnt, int n, int threshold, const int *slot, int *hot_col, int *col,
                              long long limit, int num_sms, cudaStream_t stream)
{
    hot_fill_kernel<<<(n + 255) / 256, 256, 0, stream>>>(cnt, n, threshold, slot, hot_col);
    hot_rewrite_kernel<<<num_sms * 16, 256, 0, stream>>>(col, limit, cnt, threshold, slot);
    return cudaGetLastError();
}

cudaError_t launch_hot_restore(int *col, long long limit, const int *hot_col, int num_sms, cudaStream_t stream)
{
    hot_restore_kernel<<<num_sms * 16, 256, 0, stream>>>(col, limit, hot_col);
    return cudaGetLastError();
}

cudaError_t launch_desc_offset(const Plan &pl, cudaStream_t stream)
{
    if (pl.p < 2) return cudaSuccess;
    constexpr int WPB = 8;
    const int blocks = (pl.p - 1 + WPB - 1) / WPB;
    desc_offset_kernel<WPB><<<blocks, WPB * 32, 0, stream>>>(pl.row_ptr, pl.tile_ptr, pl.desc,
                                                             pl.desc_off_ptr, pl.desc_off, pl.sigma, pl.p,
                                                             pl.bit_y, pl.bit_y + pl.bit_ss, pl.num_packet);
    return cudaGetLastError();
}

cudaError_t launch_transpose(const Plan &pl, bool r2c, cudaStream_t stream)
{
    if (pl.p < 2) return cudaSuccess;
    const int stride = pl.sigma | 1;
    const size_t smem = (size_t)OMEGA * stride * (pl.value_bytes + sizeof(int));
    if (pl.value_bytes == 8)
        transpose_kernel<double><<<pl.p - 1, 128, smem, stream>>>(pl.col, static_cast<double *>(pl.val),
                                                                  pl.tile_ptr, pl.sigma, r2c);
    else
        transpose_kernel<float><<<pl.p - 1, 128, smem, stream>>>(pl.col, static_cast<float *>(pl.val),
                                                                 pl.tile_ptr, pl.sigma, r2c);
    return cudaGetLastError();
}

cudaError_t launch_warmup(cudaStream_t stream)
{
    warmup_kernel<<<1, 32, 0, stream>>>(nullptr);
    return cudaGetLastError();
}

}  // namespace csr5
