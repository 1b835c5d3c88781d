// csr5_spmv.cuh -- the CSR5 SpMV kernels for sm_100a (included once per value type).
//
// One warp owns one omega x sigma tile (omega = 32), as in the reference
// (csr5_spmv_cuda.h:275-311), but the data path and the write-back are re-designed:
//
//  * direct-load kernel (default): one warp per tile, register-staged streaming loads (ld.global.cs) in
//    chunks sized by measurement (ChunkOf), LDG gathers of x.  On B200 it streams the banded 10M matrix at
//    1.01 of the measured HBM copy bandwidth (profiles/r01_ncu_full_c2_direct.md).
//  * TMA-staged kernel (CSR5B200_OPT_KERNEL = 2): persistent CTAs; every warp runs its own S-deep ring of
//    shared-memory slots that it fills with 1-D bulk TMA copies (cp.async.bulk -> UBLKCP) of the tile's val
//    slab, col slab and descriptor words, completing on one mbarrier per slot.  The warp that consumes a
//    slot is the warp that refills it, so there is no producer/consumer coupling between warps and no
//    empty-barrier.  Measured slower than the direct kernel (few fat warps expose the x-gather latency;
//    profiles/r01_ncu_full_c2_tma.md); with the x gathers prefetched one tile ahead (CSR5B200_OPT_KERNEL = 4)
//    it reaches 0.97-1.00 of the roofline on the banded FP64 stream (profiles/r01_tma_prefetch.txt).
//  * hot-column kernel (CSR5B200_OPT_HOT_COLUMNS): direct loads + the most referenced x entries staged in
//    shared memory by bulk TMA, for power-law matrices.
//
//  * In-lane reduction: the sigma bit flags of a lane are unpacked ONCE into a 32-bit mask, so the
//    fully unrolled loop tests compile-time bits instead of shifting a descriptor per element
//    (csr5_spmv_cuda.h:146-176).
//  * Cross-lane segmented sum: a backward segmented reduction by doubling with shuffles
//    (5 steps) over the lanes that have no row start.  It is the same sum the reference forms as
//    scan[l + seg_offset] - scan[l] + v[l] (csr5_spmv_cuda.h:25-38) without the subtraction, so it
//    has no cancellation error and a fixed association order.
//  * Write-back: every non-empty row is plain-stored exactly once by the tile in which it STARTS
//    (including a row that starts exactly on a tile boundary, which the reference routes through
//    the calibrator); only the carry-in of a tile whose first row started in an earlier tile goes
//    to calibrator[t].  A second, tiny kernel adds the non-zero calibrators with FP64/FP32
//    red.global.add.  Hence spmv() overwrites y and needs no zeroed y unless the matrix has empty
//    rows before the tail (then y is memset first).  The reference's third launch, the tail
//    kernel (csr5_spmv_cuda.h:384-419), is folded into the first: the leading warps of the grid
//    reduce the rows of the last, partial tile CSR-vector style.
//  * Sharded (multi-GPU) mode: the same kernels with every y store replicated to all GPUs (MULTI), or a
//    coalesced push pass after the SpMV; see csr5b200_spmv_scatter in include/csr5_b200.h.
#ifndef CSR5_SPMV_CUH
#define CSR5_SPMV_CUH

#include "csr5_internal.h"

namespace csr5 {

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ uint64_t l2_evict_last_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

template <typename VT>
struct SpmvArgs {
    const int *__restrict__ row_ptr;
    const int *__restrict__ col;
    const VT *__restrict__ val;
    const VT *__restrict__ x;
    VT *__restrict__ y;
    const uint32_t *__restrict__ tile_ptr;
    const uint32_t *__restrict__ desc;
    const int *__restrict__ desc_off_ptr;
    const int *__restrict__ desc_off;
    VT *__restrict__ cal;
    VT alpha;
    VT beta;             // y = alpha * A * x + beta * y (0: y is overwritten and never read)
    int m, p, bit_y, bit_all, num_packet;
    int tile_begin, tile_end;  // CSR5 tiles [tile_begin, tile_end) of [0, p - 1) this launch processes
    int tail_start;      // first row of the tail tile
    int tail_nnz_start;  // (p - 1) * omega * sigma
    int tail_warps;      // ceil((m - tail_start) / 32) when this launch also does the tail rows, else 0
    // Sharded (multi-GPU) mode (csr5b200_spmv_scatter): `y` is this rank's segment in local HBM; the
    // n_dst destination segments -- this rank's slot in each GPU's concatenated y, mapped over NVLink,
    // or ONE NVSwitch multicast address that replicates a store to all of them -- receive copies.
    int n_dst;
    int dst_multicast;   // 1: y_dst[0] is a multicast (multimem) address
    VT *y_dst[CSR5B200_MAX_SCATTER];
};

template <typename VT> __device__ __forceinline__ void multimem_store(VT *p, VT v);
template <> __device__ __forceinline__ void multimem_store<double>(double *p, double v)
{
    asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
template <> __device__ __forceinline__ void multimem_store<float>(float *p, float v)
{
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }

// All y stores of the SpMV kernels go through put_y.  MULTI = false: the plain local store.  MULTI = true
// (sharded mode, "fused" exchange): the value is stored to every destination instead -- each GPU's copy
// of the concatenated y over NVLink, or once to the NVSwitch multicast address -- so the all-gather
// traffic rides along with the tile stream.
template <bool MULTI, typename VT>
__device__ __forceinline__ void put_y(const SpmvArgs<VT> &a, int row, VT v)
{
    if constexpr (!MULTI) {
        // every row is stored exactly once per SpMV (by the tile in which it starts, or by the tail), so
        // the beta * y term is a plain read-modify-write here; carries are added afterwards
        if (a.beta != (VT)0) v = fma_t<VT>(a.beta, a.y[row], v);
        a.y[row] = v;
    } else {
        if (a.dst_multicast) {
            multimem_store<VT>(a.y_dst[0] + row, v);
        } else {
#pragma unroll
            for (int k = 0; k < CSR5B200_MAX_SCATTER; k++)
                if (k < a.n_dst) a.y_dst[k][row] = v;
        }
    }
}

// ---- small device helpers ---------------------------------------------------------------------

template <typename VT> __device__ __forceinline__ VT warp_sum_xor(VT v)
{
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) v += __shfl_xor_sync(FULL_MASK, v, w);
    return v;
}

// Lane flags from the packed descriptor words: bit i = element (lane, i) starts a row.
__device__ __forceinline__ uint32_t unpack_flags(uint32_t w0, uint32_t w1, int bit_all, int sigma)
{
    const unsigned long long word = ((unsigned long long)w0 << 32) | w1;
    uint32_t f = __brev((uint32_t)(word >> (32 - bit_all)));
    if (sigma < 32) f &= (1u << sigma) - 1u;
    return f;
}

// Where a tile's val / col / descriptor words come from.
template <typename VT>
struct GlobalTile {  // direct-load kernel: streaming loads, evict-first
    static constexpr bool kStageInRegisters = true;  // issue all val/col loads of a chunk up front
    static constexpr bool kHasX = false;
    const VT *val;
    const int *col;
    const uint32_t *desc;
    // Measured and dropped (profiles/r02_probe_cache_policy.txt): val / col through ld.global.nc.L1::no_allocate and
    // x gathers with an L2 evict_last hint move R-MAT 22 / 25 by less than 1 % -- L1TEX and L2 are both at ~80 % of
    // their peak throughput either way -- while the extra operands cost 14 registers per thread.
    __device__ __forceinline__ VT v(int i, int lane) const { return __ldcs(val + i * OMEGA + lane); }
    __device__ __forceinline__ int c(int i, int lane) const { return __ldcs(col + i * OMEGA + lane); }
    __device__ __forceinline__ uint32_t d(int k, int lane) const { return __ldg(desc + k * OMEGA + lane); }
    __device__ __forceinline__ VT xv(const VT *__restrict__ x, int c) const { return __ldg(x + c); }
};

// Direct loads as GlobalTile, but a column index with bit 31 set names a slot of the hot-column
// table staged in shared memory (Plan::hot_k) instead of an element of x.
template <typename VT>
struct HotTile {
    static constexpr bool kStageInRegisters = true;
    static constexpr bool kHasX = false;
    const VT *val;
    const int *col;
    const uint32_t *desc;
    const VT *xs;  // shared memory
    __device__ __forceinline__ VT v(int i, int lane) const { return __ldcs(val + i * OMEGA + lane); }
    __device__ __forceinline__ int c(int i, int lane) const { return __ldcs(col + i * OMEGA + lane); }
    __device__ __forceinline__ uint32_t d(int k, int lane) const { return __ldg(desc + k * OMEGA + lane); }
    __device__ __forceinline__ VT xv(const VT *__restrict__ x, int c) const
    {
        return c < 0 ? xs[c & 0x7fffffff] : __ldg(x + c);
    }
};

template <typename VT>
struct SharedTile {  // TMA-staged kernel: conflict-free LDS (consecutive lanes, consecutive words)
    static constexpr bool kStageInRegisters = false;  // val/col are one LDS away: only x needs registers
    static constexpr bool kHasX = false;
    const VT *val;
    const int *col;
    const uint32_t *desc;
    __device__ __forceinline__ VT v(int i, int lane) const { return val[i * OMEGA + lane]; }
    __device__ __forceinline__ int c(int i, int lane) const { return col[i * OMEGA + lane]; }
    __device__ __forceinline__ uint32_t d(int k, int lane) const { return desc[k * OMEGA + lane]; }
    __device__ __forceinline__ VT xv(const VT *__restrict__ x, int c) const { return __ldg(x + c); }
};

// TMA-staged kernel with x prefetch: as SharedTile, but the x values of the tile were gathered one
// tile ahead into registers (xr[i] = x[col(i, lane)]).
template <typename VT>
struct PrefetchedTile {
    static constexpr bool kStageInRegisters = false;
    static constexpr bool kHasX = true;
    const VT *val;
    const int *col;
    const uint32_t *desc;
    const VT *xr;
    __device__ __forceinline__ VT v(int i, int lane) const { return val[i * OMEGA + lane]; }
    __device__ __forceinline__ int c(int i, int lane) const { return col[i * OMEGA + lane]; }
    __device__ __forceinline__ uint32_t d(int k, int lane) const { return desc[k * OMEGA + lane]; }
    __device__ __forceinline__ VT xv(const VT *__restrict__ x, int c) const { return __ldg(x + c); }
};

// ---- one CSR5 tile (t < p - 1) ------------------------------------------------------------------
// Number of register chunks a tile's sigma elements are consumed in (NCH = 0: default rule).  Measured on
// B200 (profiles/r01_sweep_wpb_nch.txt): FP64 is fastest with <= 8 elements (96 B of stream) in flight
// per lane -- sigma 16 in two chunks streams C2 at 1.00 of the measured copy bandwidth vs 0.97 in one;
// FP32 is fastest with the whole lane (sigma <= 26) in one chunk.
template <typename VT, int SIGMA, int NCH> struct ChunkOf {
    static constexpr int AUTO = sizeof(VT) == 8 ? (SIGMA + 7) / 8 : (SIGMA <= 26 ? 1 : 2);
    static constexpr int N = NCH > 0 ? NCH : AUTO;
    static constexpr int CH = (SIGMA + N - 1) / N;
};

template <typename VT, int SIGMA, bool MULTI, typename Tile, int NCH = 0>
__device__ __forceinline__ void process_tile(const SpmvArgs<VT> &a, const Tile &tile, int t, int lane,
                                             uint32_t raw_start, uint32_t raw_stop)
{
    // Elements are consumed in register chunks so that all streaming loads and x gathers of a chunk
    // are in flight together without spilling at sigma = 32.
    constexpr int CH = ChunkOf<VT, SIGMA, NCH>::CH;
    const int row_start = (int)(raw_start & ROW_MASK);
    const int row_stop = (int)(raw_stop & ROW_MASK);
    const VT *__restrict__ x = a.x;

    if (raw_start == (uint32_t)row_stop) {
        // fast track: the whole tile lies inside one row (csr5_spmv_cuda.h:59-89)
        VT sum = 0;
#pragma unroll
        for (int c0 = 0; c0 < SIGMA; c0 += CH) {
            VT v[CH], xv[CH];
            int c[CH];
            if constexpr (Tile::kStageInRegisters) {
#pragma unroll
                for (int k = 0; k < CH; k++)
                    if (c0 + k < SIGMA) { c[k] = tile.c(c0 + k, lane); v[k] = tile.v(c0 + k, lane); }
            }
#pragma unroll
            for (int k = 0; k < CH; k++)
                if (c0 + k < SIGMA) {
                    if constexpr (Tile::kHasX) xv[k] = tile.xr[c0 + k];
                    else xv[k] = tile.xv(x, Tile::kStageInRegisters ? c[k] : tile.c(c0 + k, lane));
                }
#pragma unroll
            for (int k = 0; k < CH; k++)
                if (c0 + k < SIGMA)
                    sum = fma_t<VT>(Tile::kStageInRegisters ? v[k] : tile.v(c0 + k, lane), xv[k], sum);
        }
        sum = warp_sum_xor<VT>(sum);
        if (lane == 0) {
            const bool starts_here = (tile.d(0, 0) >> (31 - a.bit_all)) & 1u;  // raw flag of element 0
            if (starts_here) { put_y<MULTI, VT>(a, row_start, a.alpha * sum); a.cal[t] = (VT)0; }
            else a.cal[t] = a.alpha * sum;
        }
        return;
    }

    const bool dirty = raw_start >> 31;
    const uint32_t w0 = tile.d(0, lane);
    const uint32_t w1 = a.num_packet > 1 ? tile.d(1, lane) : 0u;
    const uint32_t f = unpack_flags(w0, w1, a.bit_all, SIGMA);
    const uint32_t ff = f | (lane == 0 ? 1u : 0u);  // lane 0 always opens a segment (…:138)
    int y_idx = (int)(w0 >> (32 - a.bit_y));
    const int *__restrict__ yoff = dirty ? a.desc_off + a.desc_off_ptr[t] : nullptr;
    const int Y = row_start + 1;  // rows stored by this tile are addressed relative to row_start + 1

    // `open`: the running segment began in this lane at a real row start, so it is stored
    // directly when it closes.  Lane 0's first segment (row_start itself) is handled at the end.
    bool open = (ff & 1u) && lane != 0;
    VT sum = 0, first_sum = 0;
    // (Measured and dropped, profiles/r02_ab_parked_row_stores.txt: parking the first two finished rows of a lane in
    // registers and storing them warp-wide after the loop -- fewer store instructions, i.e. fewer L1TEX wavefronts --
    // costs 8 registers and one resident CTA per SM: C2 -2..-7 %, C3 / C5 -1 %, C4 +-2 %.)
#pragma unroll
    for (int c0 = 0; c0 < SIGMA; c0 += CH) {
        VT v[CH], xv[CH];
        int c[CH];
        if constexpr (Tile::kStageInRegisters) {
#pragma unroll
            for (int k = 0; k < CH; k++)
                if (c0 + k < SIGMA) { c[k] = tile.c(c0 + k, lane); v[k] = tile.v(c0 + k, lane); }
        }
#pragma unroll
        for (int k = 0; k < CH; k++)
            if (c0 + k < SIGMA) {
                if constexpr (Tile::kHasX) xv[k] = tile.xr[c0 + k];
                else xv[k] = tile.xv(x, Tile::kStageInRegisters ? c[k] : tile.c(c0 + k, lane));
            }
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int i = c0 + k;
            if (i < SIGMA) {
                if (i > 0 && ((ff >> i) & 1u)) {
                    if (open) { put_y<MULTI, VT>(a, Y + (dirty ? yoff[y_idx] : y_idx), a.alpha * sum); y_idx++; }
                    else first_sum = sum;
                    open = true;
                    sum = 0;
                }
                sum = fma_t<VT>(Tile::kStageInRegisters ? v[k] : tile.v(i, lane), xv[k], sum);
            }
        }
    }
    if (!open) first_sum = sum;  // no row start after element 0: the whole lane is one piece
    VT last_sum = sum;

    // Cross-lane step.  carry[l] = piece of lane l that belongs to a segment opened in an earlier
    // lane (lanes whose element 0 is not a row start).  Lane l with a row start collects the
    // carries of lanes l+1 .. l+seg_offset+1, i.e. up to and including the next lane with a start.
    const VT carry = (ff & 1u) ? (VT)0 : first_sum;
    VT acc = __shfl_down_sync(FULL_MASK, carry, 1);
    if (lane == 31) acc = 0;
    const uint32_t present = __ballot_sync(FULL_MASK, ff != 0);
    const uint32_t following = lane == 31 ? 0u : present >> (lane + 1);
    const int rem = following ? __ffs(following) - 1 : 31 - lane;  // == seg_offset on lanes with a start
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const VT nb = __shfl_down_sync(FULL_MASK, acc, d);
        if (d <= rem) acc += nb;
    }
    if (ff) last_sum += acc;

    if (open) put_y<MULTI, VT>(a, Y + (dirty ? yoff[y_idx] : y_idx), a.alpha * last_sum);
    if (lane == 0) {
        const VT head = open ? first_sum : last_sum;  // lane 0's first segment = row_start's piece
        if (f & 1u) { put_y<MULTI, VT>(a, row_start, a.alpha * head); a.cal[t] = (VT)0; }  // row starts on the tile boundary
        else a.cal[t] = a.alpha * head;                                      // carry-in from earlier tiles
    }
}

// ---- rows of the tail tile: 32 rows per warp, CSR-vector per non-empty row ----------------------
template <typename VT, bool MULTI>
__device__ __forceinline__ void process_tail_rows(const SpmvArgs<VT> &a, int tw, int lane)
{
    const int r = a.tail_start + tw * 32 + lane;
    int ra = 0, rb = 0;
    bool carried = false;
    if (r < a.m) {
        ra = __ldg(a.row_ptr + r);
        rb = __ldg(a.row_ptr + r + 1);
        if (r == a.tail_start) { carried = ra != a.tail_nnz_start; ra = a.tail_nnz_start; }
    }
    VT mine = 0;
    uint32_t todo = __ballot_sync(FULL_MASK, rb > ra);
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const int ja = __shfl_sync(FULL_MASK, ra, j), jb = __shfl_sync(FULL_MASK, rb, j);
        VT s = 0;
        for (int k = ja + lane; k < jb; k += 32) s = fma_t<VT>(__ldcs(a.val + k), __ldg(a.x + __ldcs(a.col + k)), s);
        s = warp_sum_xor<VT>(s);
        if (lane == j) mine = s;
    }
    if (r < a.m) {
        if (carried) a.cal[a.p - 1] = a.alpha * mine;
        else {
            put_y<MULTI, VT>(a, r, a.alpha * mine);
            if (r == a.tail_start) a.cal[a.p - 1] = (VT)0;
        }
    }
}

// ---- direct-load kernel ---------------------------------------------------------------------------
template <typename VT, int SIGMA, int WPB, bool MULTI, int NCH = 0>
__global__ void __launch_bounds__(WPB * 32) spmv_direct_kernel(const SpmvArgs<VT> a)
{
    const int lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * WPB + (threadIdx.x >> 5);
    if (unit < a.tail_warps) { process_tail_rows<VT, MULTI>(a, (int)unit, lane); return; }
    const long long tl = a.tile_begin + (unit - a.tail_warps);
    if (tl >= a.tile_end) return;
    const int t = (int)tl;
    const size_t base = (size_t)t * (OMEGA * SIGMA);
    GlobalTile<VT> tile{a.val + base, a.col + base, a.desc + (size_t)t * OMEGA * a.num_packet};
    process_tile<VT, SIGMA, MULTI, GlobalTile<VT>, NCH>(a, tile, t, lane, __ldg(a.tile_ptr + t),
                                                        __ldg(a.tile_ptr + t + 1));
}

// ---- TMA-staged persistent kernel ---------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// 1-D bulk TMA copy global -> shared, completing (complete_tx) on an mbarrier; streamed through L2
// with an evict-first policy since every matrix byte is read once per SpMV.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar,
                                            uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

constexpr int TMA_MAX_WARPS = 16;

template <typename VT, int SIGMA>
struct TmaSlot {
    static constexpr int VAL_BYTES = OMEGA * SIGMA * (int)sizeof(VT);
    static constexpr int COL_BYTES = OMEGA * SIGMA * 4;
    static constexpr int DESC_BYTES_MAX = OMEGA * 2 * 4;
    static constexpr int BYTES = VAL_BYTES + COL_BYTES + DESC_BYTES_MAX;  // multiple of 128
};

// grid = persistent CTAs; warp gw handles tiles gw, gw + GW, gw + 2 GW, ... (GW = warps in the grid)
// PREFETCH (CSR5B200_OPT_KERNEL = 4): the x gathers of the NEXT tile are issued (from its col slab, already
// landed in the ring) before the current tile is reduced, so a warp's gather latency overlaps its own
// arithmetic instead of being exposed once per tile -- the cause of the plain ring's 87 %.
template <typename VT, int SIGMA, bool PREFETCH>
__global__ void __launch_bounds__(TMA_MAX_WARPS * 32, 1) spmv_tma_kernel(const SpmvArgs<VT> a, const int stages)
{
    extern __shared__ __align__(128) unsigned char smem[];
    using Slot = TmaSlot<VT, SIGMA>;
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const long long GW = (long long)gridDim.x * wpb;
    const long long gw = (long long)blockIdx.x * wpb + w;

    // smem: [wpb][stages] slots, then [wpb][stages] mbarriers
    unsigned char *my_slots = smem + (size_t)w * stages * Slot::BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)wpb * stages * Slot::BYTES) + w * stages;
    if (lane == 0) {
        for (int s = 0; s < stages; s++) mbar_init(smem_u32(bars + s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // tail rows first (they are few); warps stride over them
    for (long long tw = gw; tw < a.tail_warps; tw += GW) process_tail_rows<VT, false>(a, (int)tw, lane);

    const long long ntiles = a.tile_end;          // this launch: tiles [tile_begin, tile_end)
    const long long first = a.tile_begin + gw;    // this warp's first tile
    const uint32_t desc_bytes = OMEGA * 4 * a.num_packet;
    const uint32_t tx_bytes = Slot::VAL_BYTES + Slot::COL_BYTES + desc_bytes;
    const uint64_t policy = l2_evict_first_policy();

    auto issue = [&](long long t, int s) {  // lane 0 only
        unsigned char *slot = my_slots + (size_t)s * Slot::BYTES;
        const uint32_t bar = smem_u32(bars + s);
        const size_t base = (size_t)t * (OMEGA * SIGMA);
        mbar_expect_tx(bar, tx_bytes);
        tma_load_1d(smem_u32(slot), a.val + base, Slot::VAL_BYTES, bar, policy);
        tma_load_1d(smem_u32(slot + Slot::VAL_BYTES), a.col + base, Slot::COL_BYTES, bar, policy);
        tma_load_1d(smem_u32(slot + Slot::VAL_BYTES + Slot::COL_BYTES),
                    a.desc + (size_t)t * OMEGA * a.num_packet, desc_bytes, bar, policy);
    };

    // prologue: fill the ring
    if (lane == 0) {
        long long t = first;
        for (int s = 0; s < stages && t < ntiles; s++, t += GW) issue(t, s);
    }
    uint32_t raw_start = 0, raw_stop = 0;
    if (first < ntiles) { raw_start = __ldg(a.tile_ptr + first); raw_stop = __ldg(a.tile_ptr + first + 1); }

    int s = 0;
    uint32_t parity = 0;
    if constexpr (!PREFETCH) {
        for (long long t = first; t < ntiles; t += GW) {
            // tile_ptr words of the next tile are fetched one iteration ahead
            const long long tn = t + GW;
            uint32_t nstart = 0, nstop = 0;
            if (tn < ntiles) { nstart = __ldg(a.tile_ptr + tn); nstop = __ldg(a.tile_ptr + tn + 1); }

            mbar_wait(smem_u32(bars + s), parity);
            const unsigned char *slot = my_slots + (size_t)s * Slot::BYTES;
            SharedTile<VT> tile{reinterpret_cast<const VT *>(slot),
                                reinterpret_cast<const int *>(slot + Slot::VAL_BYTES),
                                reinterpret_cast<const uint32_t *>(slot + Slot::VAL_BYTES + Slot::COL_BYTES)};
            process_tile<VT, SIGMA, false>(a, tile, (int)t, lane, raw_start, raw_stop);

            // this warp is done reading slot s: refill it with the tile `stages` iterations ahead
            __syncwarp();
            const long long tf = t + (long long)stages * GW;
            if (lane == 0 && tf < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(tf, s);
            }
            raw_start = nstart;
            raw_stop = nstop;
            if (++s == stages) { s = 0; parity ^= 1u; }
        }
    } else {
        VT xc[SIGMA], xn[SIGMA];
#pragma unroll
        for (int i = 0; i < SIGMA; i++) { xc[i] = (VT)0; xn[i] = (VT)0; }
        if (first < ntiles) {  // x of this warp's first tile
            mbar_wait(smem_u32(bars), 0);
            const int *col0 = reinterpret_cast<const int *>(my_slots + Slot::VAL_BYTES);
#pragma unroll
            for (int i = 0; i < SIGMA; i++) xc[i] = __ldg(a.x + col0[i * OMEGA + lane]);
        }
        for (long long t = first; t < ntiles; t += GW) {
            const long long tn = t + GW;
            int sn = s + 1;
            uint32_t pn = parity;
            if (sn == stages) { sn = 0; pn ^= 1u; }
            uint32_t nstart = 0, nstop = 0;
            if (tn < ntiles) {
                nstart = __ldg(a.tile_ptr + tn);
                nstop = __ldg(a.tile_ptr + tn + 1);
                mbar_wait(smem_u32(bars + sn), pn);  // issued one full tile ago: normally landed
                const int *coln = reinterpret_cast<const int *>(my_slots + (size_t)sn * Slot::BYTES + Slot::VAL_BYTES);
#pragma unroll
                for (int i = 0; i < SIGMA; i++) xn[i] = __ldg(a.x + coln[i * OMEGA + lane]);
            }
            const unsigned char *slot = my_slots + (size_t)s * Slot::BYTES;
            PrefetchedTile<VT> tile{reinterpret_cast<const VT *>(slot),
                                    reinterpret_cast<const int *>(slot + Slot::VAL_BYTES),
                                    reinterpret_cast<const uint32_t *>(slot + Slot::VAL_BYTES + Slot::COL_BYTES), xc};
            process_tile<VT, SIGMA, false>(a, tile, (int)t, lane, raw_start, raw_stop);

            __syncwarp();
            const long long tf = t + (long long)stages * GW;
            if (lane == 0 && tf < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(tf, s);
            }
#pragma unroll
            for (int i = 0; i < SIGMA; i++) xc[i] = xn[i];
            raw_start = nstart;
            raw_stop = nstop;
            s = sn;
            parity = pn;
        }
    }
}

// ---- hot-column kernel (power-law matrices) ------------------------------------------------------
// The x gathers of a scale-free matrix are uncoalesced 32-byte sectors, and an SM's L1TEX pipeline
// retires about one such sector per clock -- that, not HBM, bounds the direct kernel on R-MAT
// (profiles/r01_c3_*).  Here the most referenced columns (Plan::hot_k of them, chosen at asCSR5() time)
// are served from shared memory instead: a tiny pre-kernel gathers x[hot_col[*]] into a dense array,
// every persistent CTA pulls that array into shared memory with 1-D bulk TMA copies completing on an
// mbarrier, and the tile loop reads tagged columns with LDS (a handful of bank-conflict wavefronts per
// warp instead of 32 L1TEX wavefronts).
template <typename VT>
__global__ void __launch_bounds__(256)
hot_gather_kernel(const VT *__restrict__ x, const int *__restrict__ hot_col, VT *__restrict__ hot_x, int hot_k)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < hot_k) hot_x[j] = __ldg(x + __ldg(hot_col + j));
}

constexpr int HOT_MAX_THREADS = 1024;

// (Measured and dropped, profiles/r02_probe_hot_two_ctas.txt: two 640-thread CTAs per SM with a table copy each -- 40
// resident warps at 48 registers instead of 32 at 64 -- is slower, C3 0.237 vs 0.226 ms with 8 K entries and 0.325 ms
// with 12 K: what the second table takes from L1 costs more than the extra warps bring.)
template <typename VT, int SIGMA, bool MULTI, int NCH = 0>
__global__ void __launch_bounds__(HOT_MAX_THREADS, 1)
spmv_hot_kernel(const SpmvArgs<VT> a, const VT *__restrict__ hot_x, const uint32_t hot_bytes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    VT *xs = reinterpret_cast<VT *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + hot_bytes);
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(smem_u32(bar), hot_bytes);
        const uint64_t policy = l2_evict_last_policy();  // 148 CTAs re-read the same few KB
        constexpr uint32_t PIECE = 32768;
        for (uint32_t off = 0; off < hot_bytes; off += PIECE) {
            const uint32_t n = hot_bytes - off < PIECE ? hot_bytes - off : PIECE;
            tma_load_1d(smem_u32(smem + off), reinterpret_cast<const unsigned char *>(hot_x) + off, n, smem_u32(bar),
                        policy);
        }
    }
    __syncthreads();
    mbar_wait(smem_u32(bar), 0);

    const long long GW = (long long)gridDim.x * wpb;
    const long long gw = (long long)blockIdx.x * wpb + (threadIdx.x >> 5);
    for (long long tw = gw; tw < a.tail_warps; tw += GW) process_tail_rows<VT, MULTI>(a, (int)tw, lane);
    const long long ntiles = a.tile_end;
    for (long long tl = a.tile_begin + gw; tl < ntiles; tl += GW) {
        const int t = (int)tl;
        const size_t base = (size_t)t * (OMEGA * SIGMA);
        HotTile<VT> tile{a.val + base, a.col + base, a.desc + (size_t)t * OMEGA * a.num_packet, xs};
        process_tile<VT, SIGMA, MULTI, HotTile<VT>, NCH>(a, tile, t, lane, __ldg(a.tile_ptr + t), __ldg(a.tile_ptr + t + 1));
    }
}

// ---- carries: y[row of tile t] += calibrator[t] for the tiles whose first row began earlier -----
template <typename VT>
__global__ void __launch_bounds__(256) calibrate_kernel(const SpmvArgs<VT> a, const int t_begin, const int t_end,
                                                        const int skip_row)
{
    const int t = t_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_end) return;
    const VT c = a.cal[t];
    const int row = (int)(a.tile_ptr[t] & ROW_MASK);
    if (c != (VT)0 && row != skip_row) atomicAdd(a.y + row, c);   // local HBM only, also when sharded
}

// Deterministic carry pass (CSR5B200_OPT_DETERMINISTIC): the first tile of every run of tiles that carry into the same
// row sums the run's carries in tile order and adds them to y with one plain read-modify-write -- no atomics, so the
// bits do not depend on the order in which warps retire.  Runs are long only for hub rows (R-MAT 22: <= 203 tiles).
template <typename VT>
__global__ void __launch_bounds__(256) calibrate_ordered_kernel(const SpmvArgs<VT> a, const int t_begin, const int t_end,
                                                                const int skip_row)
{
    const int t = t_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t_end) return;
    const uint32_t row = a.tile_ptr[t] & ROW_MASK;
    if ((int)row == skip_row) return;
    if (t > t_begin && (a.tile_ptr[t - 1] & ROW_MASK) == row) return;   // not the first tile of its run
    VT sum = 0;
    for (int u = t; u < t_end && (a.tile_ptr[u] & ROW_MASK) == row; u++) sum += a.cal[u];
    if (sum != (VT)0) a.y[row] += sum;
}

// Boundary pass of the row-block cut (csr5_exchange.cu): the carries a block's carry pass left out because their row
// began in an EARLIER block -- applied once every block's tiles are done, in whatever order the blocks ran.
template <typename VT>
__global__ void __launch_bounds__(256) calibrate_boundary_kernel(const SpmvArgs<VT> a, const ChunkTable tb)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.p) return;
    int lo = 0, hi = tb.n;                        // block of tile t: last c with tile_begin[c] <= t (the tail tile p - 1
    while (hi - lo > 1) {                         // belongs to the last block)
        const int mid = (lo + hi) >> 1;
        if (tb.tile_begin[mid] <= t) lo = mid; else hi = mid;
    }
    const int skip = tb.skip_row[lo];
    if (skip < 0) return;
    const VT c = a.cal[t];
    if (c != (VT)0 && (int)(a.tile_ptr[t] & ROW_MASK) == skip) atomicAdd(a.y + skip, c);
}

template <typename VT>
struct RowListArgs {
    const VT *y_local;
    VT *dst[CSR5B200_MAX_SCATTER];
    int n_dst, multicast, n;
    int rows[MAX_CHUNKS];
};

template <typename VT> __global__ void push_row_list_kernel(const RowListArgs<VT> a)
{
    const int i = threadIdx.x;
    if (i >= a.n) return;
    const int r = a.rows[i];
    const VT v = a.y_local[r];
    if (a.multicast) {
        multimem_store<VT>(a.dst[0] + r, v);
    } else {
#pragma unroll
        for (int k = 0; k < CSR5B200_MAX_SCATTER; k++)
            if (k < a.n_dst) a.dst[k][r] = v;
    }
}

// ---- sharded mode, "push" exchange: copy this rank's finished y segment to every destination ----------
// One coalesced pass after the SpMV and its carry pass: 16-byte loads from local HBM, 16-byte stores over
// NVLink to each peer or once to the NVSwitch multicast address (which replicates them in the switch).
// NVLink only ever sees wide coalesced writes, whatever the row structure of the matrix.
template <typename VT>
struct PushArgs {
    const VT *y_local;
    VT *dst[CSR5B200_MAX_SCATTER];
    int n_dst, multicast, m;
};

template <typename VT> __device__ __forceinline__ void push_store16(const PushArgs<VT> &a, size_t elem, int4 v)
{
    if (a.multicast) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.dst[0] + elem),
                     "f"(__int_as_float(v.x)), "f"(__int_as_float(v.y)), "f"(__int_as_float(v.z)),
                     "f"(__int_as_float(v.w)) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < CSR5B200_MAX_SCATTER; k++)
            if (k < a.n_dst && a.dst[k] != a.y_local) *reinterpret_cast<int4 *>(a.dst[k] + elem) = v;
    }
}

template <typename VT> __device__ __forceinline__ void push_store1(const PushArgs<VT> &a, size_t elem, VT v)
{
    if (a.multicast) {
        multimem_store<VT>(a.dst[0] + elem, v);
    } else {
#pragma unroll
        for (int k = 0; k < CSR5B200_MAX_SCATTER; k++)
            if (k < a.n_dst && a.dst[k] != a.y_local) a.dst[k][elem] = v;
    }
}

constexpr int PUSH_THREADS = 1024;   // upper bound; the launch picks the block size (csr5b200_exchange.push_threads)

template <typename VT>
__global__ void __launch_bounds__(PUSH_THREADS) push_rows_kernel(const PushArgs<VT> a)
{
    constexpr int VEC = 16 / (int)sizeof(VT);
    const size_t n = (size_t)a.m;
    // scalar head up to a 16-byte boundary (all destinations share the local segment's alignment)
    size_t head = ((16 - (reinterpret_cast<uintptr_t>(a.y_local) & 15)) & 15) / sizeof(VT);
    if (head > n) head = n;
    const size_t nvec = (n - head) / VEC;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
    if (gtid < head) push_store1<VT>(a, gtid, a.y_local[gtid]);
    const int4 *src = reinterpret_cast<const int4 *>(a.y_local + head);
    size_t i = gtid;
    for (; i + 3 * gsz < nvec; i += 4 * gsz) {   // four 16-byte loads in flight per thread before their stores
        const int4 v0 = src[i], v1 = src[i + gsz], v2 = src[i + 2 * gsz], v3 = src[i + 3 * gsz];
        push_store16<VT>(a, head + i * VEC, v0);
        push_store16<VT>(a, head + (i + gsz) * VEC, v1);
        push_store16<VT>(a, head + (i + 2 * gsz) * VEC, v2);
        push_store16<VT>(a, head + (i + 3 * gsz) * VEC, v3);
    }
    for (; i < nvec; i += gsz) push_store16<VT>(a, head + i * VEC, src[i]);
    const size_t done_elems = head + nvec * VEC;
    if (gtid < n - done_elems) push_store1<VT>(a, done_elems + gtid, a.y_local[done_elems + gtid]);
}

// sharded mode, "fused" exchange: rows no tile stores (= the empty rows) are cleared in every destination
template <typename VT>
__global__ void __launch_bounds__(256) zero_empty_rows_kernel(const SpmvArgs<VT> a)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < a.m; r += (long long)gridDim.x * blockDim.x)
        if (__ldg(a.row_ptr + r) == __ldg(a.row_ptr + r + 1)) put_y<true, VT>(a, (int)r, (VT)0);
}

// y = alpha A x + beta y: the rows no tile stores (empty rows in front of the tail) only get the beta term
template <typename VT>
__global__ void __launch_bounds__(256) scale_empty_rows_kernel(const SpmvArgs<VT> a, const int row_limit)
{
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < row_limit; r += (long long)gridDim.x * blockDim.x)
        if (__ldg(a.row_ptr + r) == __ldg(a.row_ptr + r + 1)) a.y[r] = a.beta * a.y[r];
}

// sharded mode, "fused" exchange: rows that received carries are complete in local HBM only after
// calibrate_kernel; the last tile of each such row re-sends the final value to every destination (plain
// stores -- no atomics cross NVLink).
template <typename VT>
__global__ void __launch_bounds__(256) push_carried_rows_kernel(const SpmvArgs<VT> a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.p) return;
    const uint32_t r = a.tile_ptr[t] & ROW_MASK;
    const bool last = t == a.p - 1 || (a.tile_ptr[t + 1] & ROW_MASK) != r;
    if (!last || r >= (uint32_t)a.m) return;
    const bool carried = a.cal[t] != (VT)0 || (t > 0 && (a.tile_ptr[t - 1] & ROW_MASK) == r);
    if (!carried) return;
    const VT v = a.y[r];
    if (a.dst_multicast) {
        multimem_store<VT>(a.y_dst[0] + r, v);
    } else {
#pragma unroll
        for (int k = 0; k < CSR5B200_MAX_SCATTER; k++)
            if (k < a.n_dst && a.y_dst[k] != a.y) a.y_dst[k][r] = v;
    }
}

// ---- host-side dispatch ----------------------------------------------------------------------------
template <typename VT, int SIGMA>
cudaError_t launch_sigma(const SpmvArgs<VT> &a, const SpmvTuning &tn, bool tma_ok, cudaStream_t stream,
                         int *used, int hot_k, const VT *hot_x, bool multi)
{
    const long long ntiles = (long long)a.tile_end - a.tile_begin;   // CSR5 tiles of this launch
    if (hot_k > 0 && (ntiles > 0 || a.tail_warps > 0)) {
        // column indices are tagged: only the hot-column kernel can read them
        const uint32_t hot_bytes = (uint32_t)(((size_t)hot_k * sizeof(VT) + 15) / 16 * 16);
        const size_t smem = hot_bytes + 16;
        int threads = tn.hot_threads > 0 ? tn.hot_threads : 768;
        if (threads > HOT_MAX_THREADS) threads = HOT_MAX_THREADS;
        threads = (threads + 31) / 32 * 32;
        cudaError_t e;
        auto run = [&](auto kern) -> cudaError_t {
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
                return e;
            kern<<<tn.num_sms, threads, smem, stream>>>(a, hot_x, hot_bytes);
            return cudaGetLastError();
        };
        *used = 3;
        if constexpr (SIGMA == 15 || SIGMA == 16) {
            if (!multi && tn.direct_nch == 1) return run(spmv_hot_kernel<VT, SIGMA, false, 1>);
            if (!multi && tn.direct_nch == 2) return run(spmv_hot_kernel<VT, SIGMA, false, 2>);
            if (!multi && tn.direct_nch == 3) return run(spmv_hot_kernel<VT, SIGMA, false, 3>);
            if (!multi && tn.direct_nch == 4) return run(spmv_hot_kernel<VT, SIGMA, false, 4>);
        }
        return multi ? run(spmv_hot_kernel<VT, SIGMA, true>) : run(spmv_hot_kernel<VT, SIGMA, false>);
    }
    int kernel = tn.kernel;
    // auto = direct-load: on B200 it streams C2 at 98 % of the measured HBM copy bandwidth vs 87 % for
    // the TMA-staged ring (profiles/r01_*; the ring's few, fat warps expose the x-gather latency).
    if (kernel == 0) kernel = 1;
    if ((kernel == 2 || kernel == 4) && !tma_ok) kernel = 1;
    if (ntiles <= 0 || multi) kernel = 1;  // the multi-destination (sharded) variant exists for the direct kernel
    if (kernel != 1 && a.beta != (VT)0) kernel = 1;   // beta * y is wired into the direct-load kernel's stores
    *used = kernel;

    if (kernel == 1) {
        const long long units = ntiles + a.tail_warps;
        if (units <= 0) return cudaSuccess;
        // tuning variants (CSR5B200_OPT_DIRECT_WPB / _NCH), instantiated for the sigmas of the benchmark
        // configurations only; everything else uses the default shape
        if constexpr (SIGMA == 15 || SIGMA == 16 || SIGMA == 26) {
            if (!multi && (tn.direct_wpb > 0 || tn.direct_nch > 0)) {
                const int wpb = tn.direct_wpb > 0 ? tn.direct_wpb : 4;
                const int nch = tn.direct_nch > 0 ? tn.direct_nch : ChunkOf<VT, SIGMA, 0>::N;
#define CSR5_VARIANT(W, N)                                                                                   \
    if (wpb == W && nch == N) {                                                                              \
        spmv_direct_kernel<VT, SIGMA, W, false, N><<<(unsigned)((units + W - 1) / W), W * 32, 0, stream>>>(a); \
        return cudaGetLastError();                                                                           \
    }
                CSR5_VARIANT(2, 1) CSR5_VARIANT(2, 2) CSR5_VARIANT(2, 3) CSR5_VARIANT(2, 4)
                CSR5_VARIANT(4, 1) CSR5_VARIANT(4, 2) CSR5_VARIANT(4, 3) CSR5_VARIANT(4, 4)
                CSR5_VARIANT(8, 1) CSR5_VARIANT(8, 2) CSR5_VARIANT(8, 3) CSR5_VARIANT(8, 4)
                CSR5_VARIANT(16, 2) CSR5_VARIANT(16, 4)
#undef CSR5_VARIANT
                return cudaErrorInvalidValue;
            }
        }
        constexpr int WPB = 4;
        const long long blocks = (units + WPB - 1) / WPB;
        if (multi) spmv_direct_kernel<VT, SIGMA, WPB, true><<<(unsigned)blocks, WPB * 32, 0, stream>>>(a);
        else spmv_direct_kernel<VT, SIGMA, WPB, false><<<(unsigned)blocks, WPB * 32, 0, stream>>>(a);
        return cudaGetLastError();
    }

    // Ring geometry: as many self-prefetching warps per SM as fit with `stages` slots each.
    using Slot = TmaSlot<VT, SIGMA>;
    const size_t per_slot = Slot::BYTES + 8;            // slot + its mbarrier
    const size_t smem_sm = 216 * 1024;                  // of 227 KB; leaves the per-CTA reserve
    int stages = tn.tma_stages > 0 ? tn.tma_stages : 3;
    if (stages > 8) stages = 8;
    int wps = (int)(smem_sm / (stages * per_slot));     // warps per SM
    while (wps < 4 && stages > 2) { stages--; wps = (int)(smem_sm / (stages * per_slot)); }
    if (wps < 1) wps = 1;
    if (wps > 32) wps = 32;
    int ctas_per_sm = tn.ctas_per_sm > 0 ? tn.ctas_per_sm : (wps + 7) / 8;
    int warps = tn.tma_warps > 0 ? tn.tma_warps : wps / ctas_per_sm;
    if (warps > TMA_MAX_WARPS) warps = TMA_MAX_WARPS;
    if (warps < 1) warps = 1;
    while (warps > 1 && (size_t)warps * stages * per_slot > (size_t)226 * 1024) warps--;
    const size_t smem = (size_t)warps * stages * (Slot::BYTES + 8);
    if (stages < 2 && kernel == 4) kernel = 2;  // the prefetch needs a second slot to read ahead from
    *used = kernel;
    long long grid = (long long)tn.num_sms * ctas_per_sm;
    const long long need = (ntiles + warps - 1) / warps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    auto run = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)grid, warps * 32, smem, stream>>>(a, stages);
        return cudaGetLastError();
    };
    return kernel == 4 ? run(spmv_tma_kernel<VT, SIGMA, true>) : run(spmv_tma_kernel<VT, SIGMA, false>);
}

// One group of launches of an SpMV: any subset of {prologue, tiles [tile_begin, tile_end) (+ tail rows), carry pass of the
// same tiles}.  A whole spmv() is one call with everything on; the overlapped multi-GPU exchange cuts the tiles
// into row blocks and issues the parts on different streams (csr5_exchange.cu).
template <typename VT>
cudaError_t launch_spmv_part_t(const Plan &pl, const SpmvTuning &tn, VT alpha, VT beta, VT *y, const ShardCtx *sh,
                               const SpmvCall &call, cudaStream_t stream, int *used, int *launches)
{
    if (pl.m <= 0) return cudaSuccess;
    const int n_dst = sh ? sh->n_dst : 0;
    if (n_dst < 0 || n_dst > CSR5B200_MAX_SCATTER) return cudaErrorInvalidValue;
    // exchange of the legacy sharded mode (csr5b200_spmv_scatter): fused = the SpMV kernels store to every
    // destination; push = one coalesced copy pass after the SpMV.  Auto: fused when each tile stores runs of
    // consecutive rows (no empty rows, short rows), push when the row stores are scattered.
    int exchange = sh ? sh->exchange : 0;
    if (sh && exchange == 0)
        exchange = (!pl.needs_zero_fill && pl.m > 0 && (long long)pl.nnz / pl.m <= 64) ? 1 : 2;
    const bool fused = exchange == 1;
    if (fused && beta != (VT)0) return cudaErrorNotSupported;
    cudaError_t e;
    SpmvArgs<VT> a;
    a.n_dst = n_dst;
    a.dst_multicast = sh ? sh->multicast : 0;
    for (int k = 0; k < CSR5B200_MAX_SCATTER; k++)
        a.y_dst[k] = k < n_dst ? static_cast<VT *>(sh->y_dst[k]) : nullptr;
    a.row_ptr = pl.row_ptr;
    a.col = pl.col;
    a.val = static_cast<const VT *>(pl.val);
    a.x = static_cast<const VT *>(pl.x);
    a.y = y;
    a.tile_ptr = pl.tile_ptr;
    a.desc = pl.desc;
    a.desc_off_ptr = pl.desc_off_ptr;
    a.desc_off = pl.desc_off;
    a.cal = static_cast<VT *>(pl.calibrator);
    a.alpha = alpha;
    a.beta = beta;
    a.m = pl.m;
    a.p = pl.p;
    a.bit_y = pl.bit_y;
    a.bit_all = pl.bit_y + pl.bit_ss;
    a.num_packet = pl.num_packet;
    a.tail_start = pl.tail_start;
    a.tail_nnz_start = pl.p > 0 ? (pl.p - 1) * OMEGA * pl.sigma : 0;
    const int ntiles_all = pl.p > 0 ? pl.p - 1 : 0;
    a.tile_begin = call.tile_begin < 0 ? 0 : call.tile_begin;
    a.tile_end = call.tile_end < 0 || call.tile_end > ntiles_all ? ntiles_all : call.tile_end;
    if (a.tile_begin > a.tile_end) a.tile_begin = a.tile_end;
    const bool tail = call.tail && pl.p > 0;
    a.tail_warps = tail ? (pl.m - pl.tail_start + 31) / 32 : 0;
    const int threads = 256;
    if (call.boundary) {
        if (pl.p > 0 && pl.has_carries) {
            calibrate_boundary_kernel<VT><<<(pl.p + threads - 1) / threads, threads, 0, stream>>>(a, *call.boundary);
            ++*launches;
        }
        return cudaGetLastError();
    }

    if (call.prologue) {
        // rows that no tile and no tail warp stores: the empty rows in front of the tail
        if (pl.needs_zero_fill || pl.p == 0) {
            if (fused) {
                zero_empty_rows_kernel<VT><<<tn.num_sms * 8, 256, 0, stream>>>(a);
                e = cudaGetLastError();
            } else if (beta != (VT)0) {
                scale_empty_rows_kernel<VT><<<tn.num_sms * 8, 256, 0, stream>>>(a, pl.p == 0 ? pl.m : pl.tail_start);
                e = cudaGetLastError();
            } else {
                e = cudaMemsetAsync(y, 0, (size_t)pl.m * sizeof(VT), stream);
            }
            if (e != cudaSuccess) return e;
            ++*launches;
        }
        if (pl.hot_k > 0 && pl.p > 1) {
            hot_gather_kernel<VT><<<(pl.hot_k + 255) / 256, 256, 0, stream>>>(a.x, pl.hot_col,
                                                                            static_cast<VT *>(pl.hot_x), pl.hot_k);
            if ((e = cudaGetLastError()) != cudaSuccess) return e;
            ++*launches;
        }
    }
    if (call.tiles && pl.p > 0 && (a.tile_end > a.tile_begin || a.tail_warps > 0)) {
        // bulk TMA needs 16-byte aligned global addresses; tile strides are multiples of 128 bytes
        const bool tma_ok = (reinterpret_cast<uintptr_t>(a.val) % 16 == 0) &&
                            (reinterpret_cast<uintptr_t>(a.col) % 16 == 0) &&
                            (reinterpret_cast<uintptr_t>(a.desc) % 16 == 0);
        if (tn.ev_begin && (e = cudaEventRecord(tn.ev_begin, stream)) != cudaSuccess) return e;
        switch (pl.sigma) {
#define CSR5_CASE(S) case S: e = launch_sigma<VT, S>(a, tn, tma_ok, stream, used, pl.hot_k, static_cast<const VT *>(pl.hot_x), fused); break;
            CSR5_CASE(4) CSR5_CASE(5) CSR5_CASE(6) CSR5_CASE(7) CSR5_CASE(8) CSR5_CASE(9) CSR5_CASE(10)
            CSR5_CASE(11) CSR5_CASE(12) CSR5_CASE(13) CSR5_CASE(14) CSR5_CASE(15) CSR5_CASE(16)
            CSR5_CASE(17) CSR5_CASE(18) CSR5_CASE(19) CSR5_CASE(20) CSR5_CASE(21) CSR5_CASE(22)
            CSR5_CASE(23) CSR5_CASE(24) CSR5_CASE(25) CSR5_CASE(26) CSR5_CASE(27) CSR5_CASE(28)
            CSR5_CASE(29) CSR5_CASE(30) CSR5_CASE(31) CSR5_CASE(32)
#undef CSR5_CASE
            default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
        if (tn.ev_end && (e = cudaEventRecord(tn.ev_end, stream)) != cudaSuccess) return e;
        ++*launches;
    }
    if (call.calibrate && pl.p > 0 && pl.has_carries) {   // nothing to add when every tile starts on a row boundary
        const int cb = a.tile_begin, ce = tail ? pl.p : a.tile_end;   // the tail tile's carry is calibrator[p - 1]
        if (ce > cb) {
            if (tn.deterministic)
                calibrate_ordered_kernel<VT><<<(ce - cb + threads - 1) / threads, threads, 0, stream>>>(a, cb, ce, call.skip_row);
            else
                calibrate_kernel<VT><<<(ce - cb + threads - 1) / threads, threads, 0, stream>>>(a, cb, ce, call.skip_row);   // local HBM only
            ++*launches;
        }
        if (fused) {
            push_carried_rows_kernel<VT><<<(pl.p + threads - 1) / threads, threads, 0, stream>>>(a);
            ++*launches;
        }
    }
    if (sh && !fused && call.calibrate) {
        PushArgs<VT> pa;
        pa.y_local = y;
        for (int k = 0; k < CSR5B200_MAX_SCATTER; k++) pa.dst[k] = a.y_dst[k];
        pa.n_dst = n_dst;
        pa.multicast = a.dst_multicast;
        pa.m = pl.m;
        push_rows_kernel<VT><<<tn.num_sms * 4, 256, 0, stream>>>(pa);
        ++*launches;
    }
    return cudaGetLastError();
}

template <typename VT>
cudaError_t launch_push_row_list_t(const VT *y_local, VT *const *dst, int n_dst, int multicast, const int *rows, int n,
                                   cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    RowListArgs<VT> ra;
    ra.y_local = y_local;
    for (int k = 0; k < CSR5B200_MAX_SCATTER; k++) ra.dst[k] = k < n_dst ? dst[k] : nullptr;
    ra.n_dst = n_dst;
    ra.multicast = multicast;
    ra.n = n > MAX_CHUNKS ? MAX_CHUNKS : n;
    for (int i = 0; i < ra.n; i++) ra.rows[i] = rows[i];
    push_row_list_kernel<VT><<<1, MAX_CHUNKS, 0, stream>>>(ra);
    return cudaGetLastError();
}

// Coalesced copy of `rows` finished rows of y (local HBM) to the same rows of every destination: the SM transport
// of the overlapped exchange.  `grid` CTAs only -- it shares the GPU with the SpMV of the next row block.
template <typename VT>
cudaError_t launch_push_t(const VT *y_local, VT *const *dst, int n_dst, int multicast, long long rows, int grid,
                          int threads, cudaStream_t stream)
{
    if (threads <= 0) threads = 256;
    if (threads > PUSH_THREADS) threads = PUSH_THREADS;
    threads = (threads + 31) / 32 * 32;
    if (rows <= 0) return cudaSuccess;
    PushArgs<VT> pa;
    pa.y_local = y_local;
    for (int k = 0; k < CSR5B200_MAX_SCATTER; k++) pa.dst[k] = k < n_dst ? dst[k] : nullptr;
    pa.n_dst = n_dst;
    pa.multicast = multicast;
    pa.m = (int)rows;
    push_rows_kernel<VT><<<grid, threads, 0, stream>>>(pa);
    return cudaGetLastError();
}

}  // namespace csr5

#endif
