// csr5_probe.cu -- microbenchmarks that pin the floors of the SpMV kernel on the matrix held by a handle
// (csr5b200_probe, include/csr5_b200.h).  Diagnostic only: nothing here is on the product path.
//
// The hot loop of the reference (csr5_spmv_cuda.h:144-176, `candidate` :7-23) is "stream val/col, gather x, FMA".
// On scale-free matrices the kernel is not HBM-bound but bound by the rate at which an SM's L1TEX can take
// divergent 32-byte gathers (one 128-byte line per wavefront, ~1 wavefront/clk/SM).  These probes run the SAME
// col stream with the same launch shape (one warp per tile, 4 warps per CTA, 8-element register chunks) and strip
// the kernel down to one component at a time, so the floor is measured, not estimated:
//   STREAM        val + col streamed, no x                       -> the HBM floor of the matrix stream
//   GATHER_NC     col streamed, sum += x[col] via ld.global.nc   -> the gather floor as the kernel issues it
//   GATHER_CG     the same through ld.global.cg (no L1 allocate)
//   GATHER_CA     the same through ld.global.ca
//   GATHER_ONLY   no col stream: x[hash(tile, lane, i)]          -> pure divergent-gather rate from L2
//   FMA_NOSEG     val + col + x + FMA, no descriptors / no segmented sum / one store per tile
#include "csr5_handle.h"

namespace csr5 {
namespace {

template <int KIND> __device__ __forceinline__ double load_x(const double *x, int c)
{
    if constexpr (KIND == CSR5B200_PROBE_GATHER_CG) return __ldcg(x + c);
    else if constexpr (KIND == CSR5B200_PROBE_GATHER_CA) return __ldca(x + c);
    else return __ldg(x + c);
}
template <int KIND> __device__ __forceinline__ float load_x(const float *x, int c)
{
    if constexpr (KIND == CSR5B200_PROBE_GATHER_CG) return __ldcg(x + c);
    else if constexpr (KIND == CSR5B200_PROBE_GATHER_CA) return __ldca(x + c);
    else return __ldg(x + c);
}

template <typename VT, int KIND>
__global__ void __launch_bounds__(128)
probe_kernel(const int *__restrict__ col, const VT *__restrict__ val, const VT *__restrict__ x, VT *__restrict__ out,
             const int sigma, const int ntiles, const unsigned n_mask)
{
    constexpr int CH = 8;
    const int lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= ntiles) return;
    const size_t base = (size_t)t * (OMEGA * sigma);
    VT sum = 0;
    for (int c0 = 0; c0 < sigma; c0 += CH) {
        int c[CH];
        VT v[CH], xv[CH];
#pragma unroll
        for (int k = 0; k < CH; k++) {
            c[k] = 0;
            v[k] = (VT)1;
            if (c0 + k < sigma) {
                if constexpr (KIND == CSR5B200_PROBE_GATHER_ONLY) {
                    unsigned h = (unsigned)(base + (size_t)(c0 + k) * OMEGA + lane) * 2654435761u;
                    h ^= h >> 15;
                    c[k] = (int)(h & n_mask);
                } else {
                    c[k] = __ldcs(col + base + (size_t)(c0 + k) * OMEGA + lane);
                }
                if constexpr (KIND == CSR5B200_PROBE_STREAM || KIND == CSR5B200_PROBE_FMA_NOSEG)
                    v[k] = __ldcs(val + base + (size_t)(c0 + k) * OMEGA + lane);
            }
        }
#pragma unroll
        for (int k = 0; k < CH; k++) {
            xv[k] = (VT)0;
            if (c0 + k < sigma) {
                if constexpr (KIND == CSR5B200_PROBE_STREAM) xv[k] = (VT)c[k];
                else xv[k] = load_x<KIND>(x, c[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < CH; k++)
            if (c0 + k < sigma) sum += v[k] * xv[k];
    }
#pragma unroll
    for (int w = 16; w > 0; w >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, w);
    if (lane == 0) out[t] = sum;
}

template <typename VT>
cudaError_t launch_probe(int kind, const Plan &pl, VT *out, cudaStream_t stream)
{
    const int ntiles = pl.p - 1;
    const int blocks = (ntiles + 3) / 4;
    unsigned mask = 1;
    while (mask * 2 <= (unsigned)pl.n) mask *= 2;
    mask -= 1;
    const VT *val = static_cast<const VT *>(pl.val);
    const VT *x = static_cast<const VT *>(pl.x);
    switch (kind) {
#define CSR5_PROBE(K) case K: probe_kernel<VT, K><<<blocks, 128, 0, stream>>>(pl.col, val, x, out, pl.sigma, ntiles, mask); break;
        CSR5_PROBE(CSR5B200_PROBE_STREAM)
        CSR5_PROBE(CSR5B200_PROBE_GATHER_NC)
        CSR5_PROBE(CSR5B200_PROBE_GATHER_CG)
        CSR5_PROBE(CSR5B200_PROBE_GATHER_CA)
        CSR5_PROBE(CSR5B200_PROBE_GATHER_ONLY)
        CSR5_PROBE(CSR5B200_PROBE_FMA_NOSEG)
#undef CSR5_PROBE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace
}  // namespace csr5

using namespace csr5;

extern "C" int csr5b200_probe(csr5b200_handle_t h, int kind, int repeats, float *ms_avg)
{
    if (!h || !ms_avg || repeats < 1) return CSR5B200_INVALID_ARGUMENT;
    if (h->format != CSR5B200_FORMAT_CSR5) return CSR5B200_UNSUPPORTED_CSR_SPMV;
    const Plan &pl = h->pl;
    if (pl.p < 2 || !pl.x || pl.hot_k > 0) return CSR5B200_INVALID_ARGUMENT;   // tagged columns are not indices
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaSuccess;
    auto run = [&]() -> cudaError_t {
        return pl.value_bytes == 8 ? launch_probe<double>(kind, pl, static_cast<double *>(pl.calibrator), h->stream)
                                   : launch_probe<float>(kind, pl, static_cast<float *>(pl.calibrator), h->stream);
    };
    // out = the calibrator array (p values, rewritten by every spmv)
    if ((e = cudaEventCreate(&e0)) != cudaSuccess || (e = cudaEventCreate(&e1)) != cudaSuccess) goto done;
    for (int i = 0; i < 3; i++)
        if ((e = run()) != cudaSuccess) goto done;
    if ((e = cudaEventRecord(e0, h->stream)) != cudaSuccess) goto done;
    for (int i = 0; i < repeats; i++)
        if ((e = run()) != cudaSuccess) goto done;
    if ((e = cudaEventRecord(e1, h->stream)) != cudaSuccess) goto done;
    if ((e = cudaEventSynchronize(e1)) != cudaSuccess) goto done;
    e = cudaEventElapsedTime(ms_avg, e0, e1);
    *ms_avg /= (float)repeats;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return e == cudaSuccess ? CSR5B200_SUCCESS : handle_cuda_fail(h, e);
}
