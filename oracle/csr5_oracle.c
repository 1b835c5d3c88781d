/*
 * csr5_oracle.c -- CPU restatement of the reference's CSR5_cuda algorithm (omega = 32).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library, and only as the checker.  The shipped SpMV path lives in
 * benchmark_spmv_using_csr5_b200/csrc (CUDA, sm_100a) and never calls into this file.
 *
 * Parity pinning: the reference ships no golden vectors.  This restatement
 * is pinned (tests/test_oracle.py) against
 *   (i)  the reference's own CSR5_avx2 backend compiled from /root/reference into
 *        oracle/_ref/libref_avx2.so (y, bit-exact on the reference's integer-valued input
 *        distribution, <= 1e-12 rel on real-valued inputs), and
 *   (ii) golden vectors produced by the reference's own CSR5_cuda backend (compat-patched build,
 *        oracle/build_ref_cuda.sh) run on a B200 -- tests/golden/refcuda_*.npz -- covering tile_ptr,
 *        tile_desc, the empty-row offset table, the transposed col/val arrays and y.
 *
 * Every function cites the reference file:line (relative to /root/reference/CSR5_cuda) it follows.
 * The code is a restatement, written lane-by-lane as plain loops; it is not a copy.
 *
 * Floating point: the reference is compiled by nvcc, which contracts `sum += v * x` into an FMA,
 * so the lane loops use fma()/fmaf().  The FP32 build of the reference runs the warp scan of the
 * segmented sum in FP64 (only the one-argument scan_32_shfl(double) overload matches,
 * detail/cuda/utils_cuda.h:193-216 called from csr5_spmv_cuda.h:33); restated as such.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CSR5O_OMEGA 32
#define CSR5O_MSB 0x80000000u
#define CSR5O_MASK 0x7FFFFFFFu

/* detail/common.h:13-18 */
#define CSR5O_SUCCESS 0
#define CSR5O_UNSUPPORTED_CSR5_OMEGA (-2)

/* ---------------------------------------------------------------------------------------------
 * scalars
 * ------------------------------------------------------------------------------------------- */

/* anonymouslib_cuda.h:294-318 -- auto-tuned sigma, r/s/t/u = 4/32/256/6, k = nnz / m. */
int csr5o_auto_sigma(int m, int nnz)
{
    const int k = nnz / m;
    if (k <= 4) return 4;
    if (k <= 32) return k;
    if (k <= 256) return 32;
    return 6;
}

/* anonymouslib_cuda.h:121-137 -- bit widths, packets per lane, number of tiles. */
int csr5o_layout(int sigma, int nnz, int *bit_y, int *bit_ss, int *num_packet, int *p)
{
    int base = 2, by = 1, bs = 1;
    while (base < CSR5O_OMEGA * sigma) { base *= 2; by++; }
    base = 2;
    while (base < CSR5O_OMEGA) { base *= 2; bs++; }
    *bit_y = by;
    *bit_ss = bs;
    if (by + bs > 31) return CSR5O_UNSUPPORTED_CSR5_OMEGA;
    *num_packet = (by + bs + sigma + 31) / 32;
    *p = (int)(((int64_t)nnz + (int64_t)CSR5O_OMEGA * sigma - 1) / ((int64_t)CSR5O_OMEGA * sigma));
    return CSR5O_SUCCESS;
}

/* detail/cuda/utils_cuda.h:25-53 -- number of entries of a[0..size) that are <= key. */
static int count_le(const int *a, int key, int size)
{
    int lo = 0, hi = size - 1;
    while (hi >= lo) {
        const int mid = (hi + lo) / 2;
        if (key >= a[mid]) lo = mid + 1; else hi = mid - 1;
    }
    return lo;
}

/* Flag of element (lane, i) of one tile: the (bit_all + i)-th bit, MSB first, of the lane's
 * packets laid end to end (csr5_spmv_cuda.h:137-157, format_cuda.h:203-219). */
static int flag_at(const uint32_t *tile_desc, int bit_all, int lane, int i)
{
    const int g = bit_all + i;
    return (tile_desc[(g / 32) * CSR5O_OMEGA + lane] >> (31 - g % 32)) & 1u;
}

/* ---------------------------------------------------------------------------------------------
 * CSR -> CSR5 metadata
 * ------------------------------------------------------------------------------------------- */

/* format_cuda.h:21-42 (s1: row holding each tile boundary) and 44-95 (s2: bit 31 = some row of
 * [tile_ptr[t], tile_ptr[t+1]) is empty).  tile_ptr has p + 1 entries. */
void csr5o_tile_ptr(int m, int nnz, int sigma, int p, const int *row_ptr, uint32_t *tile_ptr)
{
    for (int t = 0; t <= p; t++) {
        int64_t b = (int64_t)t * sigma * CSR5O_OMEGA;
        if (b > nnz) b = nnz;
        tile_ptr[t] = (uint32_t)(count_le(row_ptr, (int)b, m + 1) - 1);
    }
    for (int t = 0; t < p; t++) {
        const uint32_t start = tile_ptr[t] & CSR5O_MASK;
        const uint32_t stop = tile_ptr[t + 1] & CSR5O_MASK;
        if (start == stop) continue;
        for (uint32_t r = start; r < stop; r++)
            if (row_ptr[r] == row_ptr[r + 1]) { tile_ptr[t] = start | CSR5O_MSB; break; }
    }
}

/* format_cuda.h:129-159 (s1: scatter one bit per row start), 161-267 (s2: y_offset / seg_offset
 * of every lane of every non-fast-track tile t < p - 1, segment count of dirty tiles) and
 * 269-300 (s3: exclusive scan of the segment counts).
 *
 * desc        : p * 32 * num_packet words, zeroed here.
 * desc_off_ptr: p + 1 ints, zeroed here; afterwards the exclusive scan (or all zero when no tile
 *               is dirty).  *num_offsets = desc_off_ptr[p].
 *
 * Deviations from the literal reference, all on inputs where the reference itself is undefined
 * (SURVEY.md App. B): (a) a row start at nnz index >= p * 32 * sigma (trailing empty rows when
 * nnz % (32 sigma) == 0) is dropped instead of written one tile past the allocation; (b) only
 * tiles t < p - 1 get y_offset / seg_offset (the reference's rounded-up grid also touches tiles
 * p - 1 .. p + 1, which nothing ever reads); (c) desc_off_ptr[p] always receives the scan total
 * (the reference's single-block scan leaves the marker value 1 there when p % 256 == 0). */
void csr5o_tile_desc(int m, int sigma, int p, int bit_y, int bit_ss, int num_packet,
                     const int *row_ptr, const uint32_t *tile_ptr, uint32_t *desc,
                     int *desc_off_ptr, int *num_offsets)
{
    const int bit_all = bit_y + bit_ss;
    const int words = CSR5O_OMEGA * num_packet;
    memset(desc, 0, (size_t)p * words * sizeof(uint32_t));
    memset(desc_off_ptr, 0, (size_t)(p + 1) * sizeof(int));

    for (int r = 0; r < m; r++) {
        const int off = row_ptr[r];
        const int gx = off / sigma;
        const int lane = gx % CSR5O_OMEGA;
        const int t = gx / CSR5O_OMEGA;
        const int g = off % sigma + bit_all;
        if (t >= p) continue; /* deviation (a) */
        desc[(size_t)t * words + (g / 32) * CSR5O_OMEGA + lane] |= 1u << (31 - g % 32);
    }

    int any_dirty = 0;
    for (int t = 0; t < p - 1; t++) {
        uint32_t *td = desc + (size_t)t * words;
        const int dirty = (tile_ptr[t] >> 31) & 1u;
        const uint32_t start = tile_ptr[t] & CSR5O_MASK;
        const uint32_t stop = tile_ptr[t + 1] & CSR5O_MASK;
        if (start == stop) continue; /* fast-track tile keeps its raw flags */

        int segn[CSR5O_OMEGA], present[CSR5O_OMEGA];
        for (int lane = 0; lane < CSR5O_OMEGA; lane++) {
            const int f0 = flag_at(td, bit_all, lane, 0) | (lane == 0);
            int later = 0;
            for (int i = 1; i < sigma; i++) later += flag_at(td, bit_all, lane, i);
            present[lane] = f0 || later;
            const int s = later - !f0 + present[lane];
            segn[lane] = s > 0 ? s : 0;
        }
        int excl = 0;
        for (int lane = 0; lane < CSR5O_OMEGA; lane++) {
            int seg_off = 0;
            if (present[lane])
                for (int nx = lane + 1; nx < CSR5O_OMEGA && !present[nx]; nx++) seg_off++;
            const int y_off = lane ? excl - 1 : 0;
            td[lane] |= (uint32_t)y_off << (32 - bit_y);
            td[lane] |= (uint32_t)seg_off << (32 - bit_all);
            excl += segn[lane];
        }
        if (dirty) { desc_off_ptr[t] = excl; any_dirty = 1; }
    }

    if (any_dirty) {
        int run = 0;
        for (int t = 0; t < p; t++) { const int c = desc_off_ptr[t]; desc_off_ptr[t] = run; run += c; }
        desc_off_ptr[p] = run; /* deviation (c) */
    }
    *num_offsets = desc_off_ptr[p];
}

/* format_cuda.h:362-422, 472-499 -- empty-row table: for every real row-start flag of a dirty
 * tile, the index (relative to row_start + 1) of the row that starts there. */
void csr5o_desc_offset(int sigma, int p, int bit_y, int bit_ss, int num_packet,
                       const int *row_ptr, const uint32_t *tile_ptr, const uint32_t *desc,
                       const int *desc_off_ptr, int *desc_off)
{
    const int bit_all = bit_y + bit_ss;
    const int words = CSR5O_OMEGA * num_packet;
    for (int t = 0; t < p - 1; t++) {
        if (!(tile_ptr[t] >> 31)) continue;
        const int start = (int)(tile_ptr[t] & CSR5O_MASK);
        const int stop = (int)(tile_ptr[t + 1] & CSR5O_MASK);
        const uint32_t *td = desc + (size_t)t * words;
        for (int lane = 0; lane < CSR5O_OMEGA; lane++) {
            int slot = desc_off_ptr[t] + (int)(td[lane] >> (32 - bit_y));
            for (int i = 0; i < sigma; i++) {
                if (i == 0 && lane == 0) continue; /* forced flag: goes to the calibrator */
                if (!flag_at(td, bit_all, lane, i)) continue;
                const int idx = t * CSR5O_OMEGA * sigma + lane * sigma + i;
                desc_off[slot++] = count_le(row_ptr + start + 1, idx, stop - start) - 1;
            }
        }
    }
}

/* format_cuda.h:525-585 -- in-place per-tile transpose of one array of `elem` byte items.
 * r2c != 0: CSR order (lane * sigma + i) -> CSR5 order (i * 32 + lane); r2c == 0: the inverse.
 * Tiles t < p - 1 only; a tile is skipped when the RAW words tile_ptr[t] == tile_ptr[t+1]. */
void csr5o_transpose(int elem, int sigma, int nnz, const uint32_t *tile_ptr, void *data, int r2c)
{
    const int tile = CSR5O_OMEGA * sigma;
    const int p = (int)(((int64_t)nnz + tile - 1) / tile);
    char *tmp = (char *)malloc((size_t)tile * elem);
    for (int t = 0; t < p - 1; t++) {
        if (tile_ptr[t] == tile_ptr[t + 1]) continue;
        char *base = (char *)data + (size_t)t * tile * elem;
        memcpy(tmp, base, (size_t)tile * elem);
        for (int lane = 0; lane < CSR5O_OMEGA; lane++)
            for (int i = 0; i < sigma; i++) {
                const int csr = lane * sigma + i, csr5 = i * CSR5O_OMEGA + lane;
                if (r2c) memcpy(base + (size_t)csr5 * elem, tmp + (size_t)csr * elem, elem);
                else memcpy(base + (size_t)csr * elem, tmp + (size_t)csr5 * elem, elem);
            }
    }
    free(tmp);
}

/* ---------------------------------------------------------------------------------------------
 * SpMV on the CSR5 arrays (col/val already transposed), reference semantics:
 *   - y must be zero on entry; rows at tile starts, the first tail row and empty rows are only
 *     ever accumulated into / never written (csr5_spmv_cuda.h:350,377-379,418; main.cu:57);
 *   - alpha is ignored (csr5_spmv_cuda.h:22).
 * compute kernel 275-311 (fast track 59-89, normal track 91-200, segmented sum 25-38),
 * calibrate 313-382, tail 384-419.
 * ------------------------------------------------------------------------------------------- */
#define CSR5O_DEFINE_SPMV(NAME, VT, FMA)                                                          \
void NAME(int m, int sigma, int p, int bit_y, int bit_ss, int num_packet, const int *row_ptr,     \
          const int *col5, const VT *val5, const uint32_t *tile_ptr, const uint32_t *desc,        \
          const int *desc_off_ptr, const int *desc_off, const VT *x, VT *y)                       \
{                                                                                                 \
    const int bit_all = bit_y + bit_ss;                                                           \
    const int words = CSR5O_OMEGA * num_packet;                                                   \
    const int tile = CSR5O_OMEGA * sigma;                                                         \
    if (p <= 0) return;                                                                           \
    VT *cal = (VT *)calloc((size_t)p, sizeof(VT));                                                \
    for (int t = 0; t < p - 1; t++) {                                                             \
        const int *c = col5 + (size_t)t * tile;                                                   \
        const VT *v = val5 + (size_t)t * tile;                                                    \
        const uint32_t raw = tile_ptr[t];                                                         \
        const int row_start = (int)(raw & CSR5O_MASK);                                            \
        const int row_stop = (int)(tile_ptr[t + 1] & CSR5O_MASK);                                 \
        if ((uint32_t)raw == (uint32_t)row_stop) { /* fast track: raw word, MSB clear */          \
            VT lane_sum[CSR5O_OMEGA];                                                             \
            for (int lane = 0; lane < CSR5O_OMEGA; lane++) {                                      \
                VT s = 0;                                                                         \
                for (int i = 0; i < sigma; i++)                                                   \
                    s = FMA(v[i * CSR5O_OMEGA + lane], x[c[i * CSR5O_OMEGA + lane]], s);          \
                lane_sum[lane] = s;                                                               \
            }                                                                                     \
            for (int w = CSR5O_OMEGA / 2; w > 0; w >>= 1) { /* xor butterfly, utils_cuda.h:99 */  \
                VT nx[CSR5O_OMEGA];                                                               \
                for (int lane = 0; lane < CSR5O_OMEGA; lane++)                                    \
                    nx[lane] = lane_sum[lane] + lane_sum[lane ^ w];                               \
                memcpy(lane_sum, nx, sizeof(nx));                                                 \
            }                                                                                     \
            cal[t] = lane_sum[0];                                                                 \
            continue;                                                                             \
        }                                                                                         \
        const int dirty = (raw >> 31) & 1u;                                                       \
        const uint32_t *td = desc + (size_t)t * words;                                            \
        const int base = dirty ? desc_off_ptr[t] : 0;                                             \
        VT *Y = y + row_start + 1;                                                                \
        VT first_sum[CSR5O_OMEGA], last_sum[CSR5O_OMEGA];                                         \
        int start[CSR5O_OMEGA], stop[CSR5O_OMEGA], direct[CSR5O_OMEGA], y_off[CSR5O_OMEGA];       \
        int seg_off[CSR5O_OMEGA];                                                                 \
        for (int lane = 0; lane < CSR5O_OMEGA; lane++) {                                          \
            const uint32_t w0 = td[lane];                                                         \
            int yo = (int)(w0 >> (32 - bit_y));                                                   \
            seg_off[lane] = (int)((w0 << bit_y) >> (32 - bit_ss));                                \
            const int f0 = flag_at(td, bit_all, lane, 0) | (lane == 0);                           \
            int dir = f0 && lane != 0, st = 0;                                                    \
            VT fs = 0, s = FMA(v[lane], x[c[lane]], (VT)0);                                       \
            for (int i = 1; i < sigma; i++) {                                                     \
                if (flag_at(td, bit_all, lane, i)) {                                              \
                    if (dir) { Y[dirty ? desc_off[base + yo] : yo] = s; yo++; }                   \
                    else fs = s;                                                                  \
                    dir = 1; s = 0; st++;                                                         \
                }                                                                                 \
                s = FMA(v[i * CSR5O_OMEGA + lane], x[c[i * CSR5O_OMEGA + lane]], s);              \
            }                                                                                     \
            first_sum[lane] = dir ? fs : s;                                                       \
            last_sum[lane] = s;                                                                   \
            start[lane] = !f0; stop[lane] = st; direct[lane] = dir; y_off[lane] = yo;             \
        }                                                                                         \
        /* segmented sum: shift down by one lane, inclusive scan (in double), then               \
         * scan[lane + seg_off] - scan[lane] + shifted[lane]  (csr5_spmv_cuda.h:25-38) */         \
        double sh[CSR5O_OMEGA], sc[CSR5O_OMEGA];                                                  \
        for (int lane = 0; lane < CSR5O_OMEGA; lane++) {                                          \
            const int nx = lane + 1;                                                              \
            const VT vv = (nx < CSR5O_OMEGA && start[nx]) ? first_sum[nx] : (VT)0;                \
            sh[lane] = (double)vv;                                                                \
        }                                                                                         \
        memcpy(sc, sh, sizeof(sc));                                                               \
        for (int d = 1; d < CSR5O_OMEGA; d <<= 1) { /* Hillis-Steele, utils_cuda.h:193-216 */     \
            double nx[CSR5O_OMEGA];                                                               \
            for (int lane = 0; lane < CSR5O_OMEGA; lane++)                                        \
                nx[lane] = lane >= d ? sc[lane] + sc[lane - d] : sc[lane];                        \
            memcpy(sc, nx, sizeof(nx));                                                           \
        }                                                                                         \
        for (int lane = 0; lane < CSR5O_OMEGA; lane++) {                                          \
            int src = lane + seg_off[lane];                                                       \
            if (src >= CSR5O_OMEGA) src = lane; /* shfl_down out of range returns own value */    \
            const VT carry = (VT)(sc[src] - sc[lane] + sh[lane]);                                 \
            if (start[lane] <= stop[lane]) last_sum[lane] += carry;                               \
            if (direct[lane]) Y[dirty ? desc_off[base + y_off[lane]] : y_off[lane]] = last_sum[lane]; \
        }                                                                                         \
        cal[t] = direct[0] ? first_sum[0] : last_sum[0];                                          \
    }                                                                                             \
    for (int t = 0; t < p - 1; t++) y[tile_ptr[t] & CSR5O_MASK] += cal[t];                        \
    /* tail tile, CSR-vector: 32 strided partial sums, xor-butterfly reduce */                    \
    const int tail_start = (int)(tile_ptr[p - 1] & CSR5O_MASK);                                   \
    for (int r = tail_start; r < m; r++) {                                                        \
        const int a = (r == tail_start) ? (p - 1) * tile : row_ptr[r];                            \
        const int b = row_ptr[r + 1];                                                             \
        VT lane_sum[CSR5O_OMEGA];                                                                 \
        for (int lane = 0; lane < CSR5O_OMEGA; lane++) {                                          \
            VT s = 0;                                                                             \
            for (int j = a + lane; j < b; j += CSR5O_OMEGA) s = FMA(val5[j], x[col5[j]], s);      \
            lane_sum[lane] = s;                                                                   \
        }                                                                                         \
        for (int w = CSR5O_OMEGA / 2; w > 0; w >>= 1) {                                           \
            VT nx[CSR5O_OMEGA];                                                                   \
            for (int lane = 0; lane < CSR5O_OMEGA; lane++)                                        \
                nx[lane] = lane_sum[lane] + lane_sum[lane ^ w];                                   \
            memcpy(lane_sum, nx, sizeof(nx));                                                     \
        }                                                                                         \
        y[r] = (r == tail_start) ? y[r] + lane_sum[0] : lane_sum[0];                              \
    }                                                                                             \
    free(cal);                                                                                    \
}

CSR5O_DEFINE_SPMV(csr5o_spmv_f64, double, fma)
CSR5O_DEFINE_SPMV(csr5o_spmv_f32, float, fmaf)

/* ---------------------------------------------------------------------------------------------
 * One-call convenience: CSR in (not modified), reference-semantics y out (y zeroed here, as the
 * reference's caller does, main.cu:57).  sigma <= 0 selects the auto rule.  Returns 0 or a
 * reference error code.
 * ------------------------------------------------------------------------------------------- */
#define CSR5O_DEFINE_FULL(NAME, VT, SPMV)                                                         \
int NAME(int m, int n, int nnz, int sigma, const int *row_ptr, const int *col, const VT *val,     \
         const VT *x, VT *y)                                                                      \
{                                                                                                 \
    (void)n;                                                                                      \
    memset(y, 0, (size_t)m * sizeof(VT));                                                         \
    if (nnz <= 0 || m <= 0) return CSR5O_SUCCESS;                                                 \
    if (sigma <= 0) sigma = csr5o_auto_sigma(m, nnz);                                             \
    int bit_y, bit_ss, np, p;                                                                     \
    const int err = csr5o_layout(sigma, nnz, &bit_y, &bit_ss, &np, &p);                           \
    if (err) return err;                                                                          \
    uint32_t *tp = (uint32_t *)malloc((size_t)(p + 1) * sizeof(uint32_t));                        \
    uint32_t *desc = (uint32_t *)malloc((size_t)p * CSR5O_OMEGA * np * sizeof(uint32_t));         \
    int *dop = (int *)malloc((size_t)(p + 1) * sizeof(int));                                      \
    int *c5 = (int *)malloc((size_t)nnz * sizeof(int));                                           \
    VT *v5 = (VT *)malloc((size_t)nnz * sizeof(VT));                                              \
    memcpy(c5, col, (size_t)nnz * sizeof(int));                                                   \
    memcpy(v5, val, (size_t)nnz * sizeof(VT));                                                    \
    int num_offsets = 0;                                                                          \
    csr5o_tile_ptr(m, nnz, sigma, p, row_ptr, tp);                                                \
    csr5o_tile_desc(m, sigma, p, bit_y, bit_ss, np, row_ptr, tp, desc, dop, &num_offsets);        \
    int *doff = (int *)malloc((size_t)(num_offsets > 0 ? num_offsets : 1) * sizeof(int));         \
    if (num_offsets) csr5o_desc_offset(sigma, p, bit_y, bit_ss, np, row_ptr, tp, desc, dop, doff);\
    csr5o_transpose((int)sizeof(int), sigma, nnz, tp, c5, 1);                                     \
    csr5o_transpose((int)sizeof(VT), sigma, nnz, tp, v5, 1);                                      \
    SPMV(m, sigma, p, bit_y, bit_ss, np, row_ptr, c5, v5, tp, desc, dop, doff, x, y);             \
    free(tp); free(desc); free(dop); free(c5); free(v5); free(doff);                              \
    return CSR5O_SUCCESS;                                                                         \
}

CSR5O_DEFINE_FULL(csr5o_csr5_spmv_f64, double, csr5o_spmv_f64)
CSR5O_DEFINE_FULL(csr5o_csr5_spmv_f32, float, csr5o_spmv_f32)

/* main.cu:336-350 -- the reference's own pass/fail yardstick: scalar CSR loop on one core,
 * sum += x[col] * val * alpha. */
void csr5o_csr_spmv_f64(int m, const int *row_ptr, const int *col, const double *val,
                        const double *x, double alpha, double *y)
{
    for (int i = 0; i < m; i++) {
        double sum = 0;
        for (int j = row_ptr[i]; j < row_ptr[i + 1]; j++) sum += x[col[j]] * val[j] * alpha;
        y[i] = sum;
    }
}

void csr5o_csr_spmv_f32(int m, const int *row_ptr, const int *col, const float *val,
                        const float *x, float alpha, float *y)
{
    for (int i = 0; i < m; i++) {
        float sum = 0;
        for (int j = row_ptr[i]; j < row_ptr[i + 1]; j++) sum += x[col[j]] * val[j] * alpha;
        y[i] = sum;
    }
}

/* y = alpha * (A x) + beta * y -- the scalar statement of the form the reference's spmv() stubs out (the
 * commented-out `beta` argument at anonymouslib_cuda.h:281, next to the alpha its kernels ignore,
 * csr5_spmv_cuda.h:22).  Row sums as in main.cu:343-349. */
void csr5o_csr_axpby_f64(int m, const int *row_ptr, const int *col, const double *val,
                         const double *x, double alpha, double beta, double *y)
{
    for (int i = 0; i < m; i++) {
        double sum = 0;
        for (int j = row_ptr[i]; j < row_ptr[i + 1]; j++) sum += x[col[j]] * val[j];
        y[i] = alpha * sum + beta * y[i];
    }
}

void csr5o_csr_axpby_f32(int m, const int *row_ptr, const int *col, const float *val,
                         const float *x, float alpha, float beta, float *y)
{
    for (int i = 0; i < m; i++) {
        float sum = 0;
        for (int j = row_ptr[i]; j < row_ptr[i + 1]; j++) sum += x[col[j]] * val[j];
        y[i] = alpha * sum + beta * y[i];
    }
}

/* FP32 inputs, FP64 accumulation: error yardstick for the FP32 configuration. */
void csr5o_csr_spmv_f32_acc64(int m, const int *row_ptr, const int *col, const float *val,
                              const float *x, double *y)
{
    for (int i = 0; i < m; i++) {
        double sum = 0;
        for (int j = row_ptr[i]; j < row_ptr[i + 1]; j++) sum += (double)x[col[j]] * (double)val[j];
        y[i] = sum;
    }
}

/* OpenMP-free multi-row scalar CSR used as the FP32 CPU baseline is in ref_avx2_driver.cpp. */
