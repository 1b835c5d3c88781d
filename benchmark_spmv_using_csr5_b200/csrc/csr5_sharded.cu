// csr5_sharded.cu -- single-process host API of the row-range sharded CSR5 SpMV (include/csr5_b200_sharded.h).
//
// No reference counterpart (the reference is single-device).  Every shard is an ordinary
// csr5b200 handle on its own device; the step is csr5b200_spmv_allgather (csr5_exchange.cu) on every shard, issued
// by one worker thread per shard so that the launch streams of the GPUs fill concurrently (a step is ~100 CUDA
// calls per shard with the copy-engine transport, next to well under a millisecond of device time).
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/csr5_b200_sharded.h"
#include "csr5_handle.h"

using namespace csr5;

namespace {

struct Shard {
    int dev = 0;
    csr5b200_handle_t h = nullptr;
    cudaStream_t stream = nullptr;
    int *row_ptr = nullptr, *col = nullptr;
    void *val = nullptr, *x = nullptr, *ybuf = nullptr;
    uint32_t *flags = nullptr;
    cudaEvent_t ev_done = nullptr;
    long long row_begin = 0, rows = 0;
    int nnz = 0;
    int err = 0;
};

// One thread per shard, all executing the same job on their own shard index; run() returns when all are done.
class Workers {
public:
    explicit Workers(int n) : n_(n)
    {
        for (int i = 0; i < n_; i++) threads_.emplace_back([this, i] { loop(i); });
    }
    ~Workers()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void run(const std::function<void(int)> &job)
    {
        std::unique_lock<std::mutex> lk(mu_);
        job_ = &job;
        pending_ = n_;
        gen_++;
        cv_.notify_all();
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int i)
    {
        unsigned long long seen = 0;
        for (;;) {
            const std::function<void(int)> *job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                job = job_;
            }
            (*job)(i);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *job_ = nullptr;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

}  // namespace

struct csr5b200_sharded_s {
    int G = 0, vb = 8;
    int m = 0, n = 0, nnz = 0;
    size_t stride = 0;           // values between the two y buffers
    int parity = 0, last = 0;    // buffer the next step writes / the last step wrote
    int transport = CSR5B200_TRANSPORT_AUTO, chunks = 0, push_ctas = 0, barrier = CSR5B200_BARRIER_AUTO, timeout_ms = 0;
    bool shared_device = false, have_matrix = false, csr5 = false;
    double row_cost = 0.0;       // > 0: shards minimise max(nnz, row_cost * rows) instead of balancing nnz
    std::vector<Shard> sh;
    std::vector<long long> bounds;
    Workers *workers = nullptr;
    int last_cuda_error = 0;
};

namespace {

int first_error(csr5b200_sharded_t s)
{
    for (auto &x : s->sh)
        if (x.err) return x.err;
    return CSR5B200_SUCCESS;
}

// f(shard index) on every shard, each on its own worker thread with its device current.
void for_each_shard(csr5b200_sharded_t s, const std::function<int(int)> &f)
{
    auto job = [&](int i) {
        Shard &x = s->sh[i];
        cudaError_t e = cudaSetDevice(x.dev);
        x.err = e != cudaSuccess ? CSR5B200_CUDA_ERROR : f(i);
    };
    if (s->G == 1 || !s->workers) {
        for (int i = 0; i < s->G; i++) job(i);
    } else {
        s->workers->run(job);
    }
}

int cu(csr5b200_sharded_t s, cudaError_t e)
{
    if (e == cudaSuccess) return CSR5B200_SUCCESS;
    s->last_cuda_error = (int)e;
    return CSR5B200_CUDA_ERROR;
}

#define CUS(call)                                    \
    do {                                             \
        const int c__ = cu(s, (call));               \
        if (c__) return c__;                         \
    } while (0)

int count_le_host(const int *a, int n, long long key)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

void free_matrix(csr5b200_sharded_t s)
{
    for (auto &x : s->sh) {
        cudaSetDevice(x.dev);
        if (x.h) csr5b200_destroy(x.h);   // back to CSR; releases the CSR5 arrays
        cudaFree(x.row_ptr);
        cudaFree(x.col);
        cudaFree(x.val);
        cudaFree(x.x);
        cudaFree(x.ybuf);
        x.row_ptr = x.col = nullptr;
        x.val = x.x = x.ybuf = nullptr;
    }
    s->have_matrix = s->csr5 = false;
}

int use_flags(csr5b200_sharded_t s)
{
    if (s->G == 1) return 0;
    if (s->barrier == CSR5B200_BARRIER_FLAGS) return 1;
    if (s->barrier == CSR5B200_BARRIER_EVENTS) return 0;
    return s->shared_device ? 0 : 1;
}

}  // namespace

extern "C" {

int csr5b200_sharded_create(int n_shards, const int *devices, int value_bytes, csr5b200_sharded_t *out)
{
    if (!out || !devices || n_shards < 1 || n_shards > CSR5B200_MAX_SCATTER) return CSR5B200_INVALID_ARGUMENT;
    if (value_bytes != 4 && value_bytes != 8) return CSR5B200_UNSUPPORTED_VALUE_TYPE;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) return CSR5B200_CUDA_ERROR;
    for (int i = 0; i < n_shards; i++)
        if (devices[i] < 0 || devices[i] >= ndev) return CSR5B200_INVALID_ARGUMENT;
    csr5b200_sharded_t s = new (std::nothrow) csr5b200_sharded_s();
    if (!s) return CSR5B200_INVALID_ARGUMENT;
    s->G = n_shards;
    s->vb = value_bytes;
    s->sh.resize(n_shards);
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    for (int i = 0; i < n_shards; i++) {
        s->sh[i].dev = devices[i];
        for (int j = 0; j < i; j++) s->shared_device |= devices[j] == devices[i];
    }
    int err = CSR5B200_SUCCESS;
    for (int i = 0; i < n_shards && !err; i++) {
        Shard &x = s->sh[i];
        if ((err = cu(s, cudaSetDevice(x.dev)))) break;
        for (int j = 0; j < n_shards; j++) {
            if (devices[j] == x.dev) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, x.dev, devices[j]);
            if (!can) { err = CSR5B200_INVALID_ARGUMENT; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { err = cu(s, e); break; }
            cudaGetLastError();
        }
        if (err) break;
        if ((err = cu(s, cudaStreamCreateWithFlags(&x.stream, cudaStreamNonBlocking)))) break;
        if ((err = cu(s, cudaEventCreateWithFlags(&x.ev_done, cudaEventDisableTiming)))) break;
        if ((err = cu(s, cudaMalloc(&x.flags, 64 * sizeof(uint32_t))))) break;
        if ((err = cu(s, cudaMemset(x.flags, 0, 64 * sizeof(uint32_t))))) break;
    }
    cudaSetDevice(prev_dev);
    if (err) {
        csr5b200_sharded_destroy(s);
        return err;
    }
    if (n_shards > 1) s->workers = new Workers(n_shards);
    *out = s;
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_input_csr_host(csr5b200_sharded_t s, int m, int n, int nnz, const int *row_ptr, const int *col,
                                    const void *val)
{
    if (!s || m < 0 || n < 0 || nnz < 0 || !row_ptr || (nnz > 0 && (!col || !val))) return CSR5B200_INVALID_ARGUMENT;
    if (row_ptr[0] != 0 || row_ptr[m] != nnz) return CSR5B200_INVALID_ARGUMENT;
    free_matrix(s);
    s->m = m;
    s->n = n;
    s->nnz = nnz;
    const int G = s->G;
    // shard g starts at the row that holds nnz index g * nnz / G, the LAST such row on ties
    // (format_cuda.h:31-41 / utils_cuda.h:25-53 applied to shard boundaries)
    s->bounds.assign(G + 1, 0);
    if (s->row_cost > 0.0) {
        // minimise max over shards of max(nnz, row_cost * rows): smallest T for which a left-to-right sweep that closes
        // a range as late as both limits allow covers all rows with G ranges (the rule of sharded.row_partition)
        const double rc = s->row_cost;
        auto sweep = [&](long long T, std::vector<long long> *out) {
            long long r = 0;
            const long long max_rows = (long long)((double)T / rc);
            for (int g = 0; g < G; g++) {
                if (r < m) {
                    long long r1 = (long long)count_le_host(row_ptr, m + 1, (long long)row_ptr[r] + T) - 1;
                    if (r1 > r + max_rows) r1 = r + max_rows;
                    if (r1 > m) r1 = m;
                    if (r1 <= r) return false;   // one row alone exceeds T
                    r = r1;
                }
                if (out) (*out)[g + 1] = r;
            }
            return r >= m;
        };
        long long lo = 1, hi = (long long)((double)nnz + rc * m) + 1;
        while (lo < hi) {
            const long long mid = lo + (hi - lo) / 2;
            if (sweep(mid, nullptr)) hi = mid; else lo = mid + 1;
        }
        sweep(lo, &s->bounds);
    } else {
        for (int g = 1; g < G; g++) {
            long long b = (long long)count_le_host(row_ptr, m + 1, (long long)nnz * g / G) - 1;
            if (b < s->bounds[g - 1]) b = s->bounds[g - 1];
            if (b > m) b = m;
            s->bounds[g] = b;
        }
    }
    s->bounds[G] = m;
    s->stride = ((size_t)m + 31) / 32 * 32;
    const size_t vb = (size_t)s->vb;
    for_each_shard(s, [&](int i) -> int {
        Shard &x = s->sh[i];
        x.row_begin = s->bounds[i];
        x.rows = s->bounds[i + 1] - s->bounds[i];
        const int a = row_ptr[x.row_begin], b = row_ptr[x.row_begin + x.rows];
        x.nnz = b - a;
        std::vector<int> rp((size_t)x.rows + 1);
        for (long long r = 0; r <= x.rows; r++) rp[r] = row_ptr[x.row_begin + r] - a;
        cudaError_t e;
        if ((e = cudaMalloc(&x.row_ptr, ((size_t)x.rows + 1) * sizeof(int))) != cudaSuccess ||
            (e = cudaMalloc(&x.col, (size_t)(x.nnz > 0 ? x.nnz : 1) * sizeof(int))) != cudaSuccess ||
            (e = cudaMalloc(&x.val, (size_t)(x.nnz > 0 ? x.nnz : 1) * vb)) != cudaSuccess ||
            (e = cudaMalloc(&x.ybuf, 2 * (s->stride ? s->stride : 32) * vb)) != cudaSuccess ||
            (e = cudaMemset(x.ybuf, 0, 2 * (s->stride ? s->stride : 32) * vb)) != cudaSuccess ||
            (e = cudaMemcpy(x.row_ptr, rp.data(), rp.size() * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(x.col, col + a, (size_t)x.nnz * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMemcpy(x.val, static_cast<const char *>(val) + (size_t)a * vb, (size_t)x.nnz * vb,
                            cudaMemcpyHostToDevice)) != cudaSuccess)
            return cu(s, e);
        if (x.h) { csr5b200_free(x.h); x.h = nullptr; }
        int err = csr5b200_create((int)x.rows, n, s->vb, &x.h);
        if (err) return err;
        if ((err = csr5b200_set_stream(x.h, x.stream))) return err;
        if ((err = csr5b200_input_csr(x.h, x.nnz, x.row_ptr, x.col, x.val))) return err;
        return csr5b200_set_sigma(x.h, CSR5B200_AUTO_TUNED_SIGMA);
    });
    const int err = first_error(s);
    if (!err) s->have_matrix = true;
    s->parity = s->last = 0;
    return err;
}

int csr5b200_sharded_set_partition(csr5b200_sharded_t s, double row_cost)
{
    if (!s || !(row_cost >= 0.0)) return CSR5B200_INVALID_ARGUMENT;
    s->row_cost = row_cost;
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_set_sigma(csr5b200_sharded_t s, int sigma)
{
    if (!s || !s->have_matrix || s->csr5) return CSR5B200_INVALID_ARGUMENT;
    for (auto &x : s->sh) {
        const int err = csr5b200_set_sigma(x.h, sigma);
        if (err) return err;
    }
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_set_option(csr5b200_sharded_t s, int option, int value)
{
    if (!s || !s->have_matrix) return CSR5B200_INVALID_ARGUMENT;
    for (auto &x : s->sh) {
        const int err = csr5b200_set_option(x.h, option, value);
        if (err) return err;
    }
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_set_exchange(csr5b200_sharded_t s, int transport, int chunks, int push_ctas, int barrier,
                                  int timeout_ms)
{
    if (!s || transport < 0 || transport > CSR5B200_TRANSPORT_NONE || transport == CSR5B200_TRANSPORT_SM_MULTICAST ||
        barrier < 0 || barrier > CSR5B200_BARRIER_EVENTS || chunks < 0 || push_ctas < 0)
        return CSR5B200_INVALID_ARGUMENT;
    s->transport = transport;
    s->chunks = chunks;
    s->push_ctas = push_ctas;
    s->barrier = barrier;
    s->timeout_ms = timeout_ms;
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_set_x_host(csr5b200_sharded_t s, const void *x_host)
{
    if (!s || !s->have_matrix || (!x_host && s->n > 0)) return CSR5B200_INVALID_ARGUMENT;
    const size_t bytes = (size_t)s->n * s->vb;
    for_each_shard(s, [&](int i) -> int {
        Shard &x = s->sh[i];
        cudaError_t e;
        if (!x.x && (e = cudaMalloc(&x.x, bytes ? bytes : 1)) != cudaSuccess) return cu(s, e);
        // ordered after the steps already enqueued on this shard (they may still read the old x)
        if ((e = cudaMemcpyAsync(x.x, x_host, bytes, cudaMemcpyHostToDevice, x.stream)) != cudaSuccess) return cu(s, e);
        if ((e = cudaStreamSynchronize(x.stream)) != cudaSuccess) return cu(s, e);
        return csr5b200_set_x(x.h, x.x);
    });
    return first_error(s);
}

int csr5b200_sharded_as_csr5(csr5b200_sharded_t s)
{
    if (!s || !s->have_matrix) return CSR5B200_INVALID_ARGUMENT;
    for_each_shard(s, [&](int i) -> int { return csr5b200_as_csr5(s->sh[i].h); });
    const int err = first_error(s);
    if (!err) s->csr5 = true;
    return err;
}

static int sharded_step(csr5b200_sharded_t s, double alpha, double beta)
{
    const int flags = use_flags(s);
    const int b = s->parity;
    const size_t vb = (size_t)s->vb;
    for_each_shard(s, [&](int i) -> int {
        Shard &x = s->sh[i];
        csr5b200_exchange ex = {};
        ex.rank = i;
        ex.world = s->G;
        for (int k = 0; k < s->G; k++) {
            ex.y_full[k] = static_cast<char *>(s->sh[k].ybuf) + (size_t)b * s->stride * vb;
            ex.flags[k] = flags ? s->sh[k].flags : nullptr;
        }
        ex.y_multicast = nullptr;
        ex.row_begin = x.row_begin;
        ex.chunks = s->chunks;
        ex.transport = s->transport;
        ex.entry_barrier = 0;   // the two y buffers alternate: nobody can still be reading the one written now
        ex.push_ctas = s->push_ctas;
        ex.timeout_ms = s->timeout_ms;
        if (beta != 0.0 && s->last != b) {
            // beta * y refers to the y of the previous step, which lives in the other buffer
            const char *prev = static_cast<const char *>(x.ybuf) + ((size_t)s->last * s->stride + x.row_begin) * vb;
            char *cur = static_cast<char *>(x.ybuf) + ((size_t)b * s->stride + x.row_begin) * vb;
            const cudaError_t e = cudaMemcpyAsync(cur, prev, (size_t)x.rows * vb, cudaMemcpyDeviceToDevice, x.stream);
            if (e != cudaSuccess) return cu(s, e);
        }
        const int err = csr5b200_spmv_allgather(x.h, alpha, beta, &ex);
        if (err) return err;
        if (!flags && s->G > 1) {
            const cudaError_t e = cudaEventRecord(x.ev_done, x.stream);
            if (e != cudaSuccess) return cu(s, e);
        }
        return CSR5B200_SUCCESS;
    });
    int err = first_error(s);
    if (err) return err;
    if (!flags && s->G > 1) {
        // host-enqueued barrier: all events are recorded (the workers have returned); every stream waits for all
        for (int i = 0; i < s->G; i++) {
            CUS(cudaSetDevice(s->sh[i].dev));
            for (int k = 0; k < s->G; k++)
                if (k != i) CUS(cudaStreamWaitEvent(s->sh[i].stream, s->sh[k].ev_done, 0));
        }
    }
    s->last = b;
    s->parity = b ^ 1;
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_spmv(csr5b200_sharded_t s, double alpha, double beta)
{
    if (!s || !s->csr5) return s && s->have_matrix ? CSR5B200_UNSUPPORTED_CSR_SPMV : CSR5B200_INVALID_ARGUMENT;
    int dev = 0;
    cudaGetDevice(&dev);
    const int err = sharded_step(s, alpha, beta);
    cudaSetDevice(dev);
    return err;
}

int csr5b200_sharded_iterate(csr5b200_sharded_t s, int steps, double alpha)
{
    if (!s || !s->csr5) return s && s->have_matrix ? CSR5B200_UNSUPPORTED_CSR_SPMV : CSR5B200_INVALID_ARGUMENT;
    if (s->m != s->n || steps < 0) return CSR5B200_INVALID_ARGUMENT;
    int dev = 0;
    cudaGetDevice(&dev);
    int err = CSR5B200_SUCCESS;
    for (int it = 0; it < steps && !err; it++) {
        err = sharded_step(s, alpha, 0.0);
        if (err) break;
        // the gathered y of this step is the x of the next one, on every device
        for (auto &x : s->sh) {
            err = csr5b200_set_x(x.h, static_cast<char *>(x.ybuf) + (size_t)s->last * s->stride * s->vb);
            if (err) break;
        }
    }
    if (!err && steps > 0) {
        // leave x in the shards' own x buffers: the y buffer it lives in is rewritten by the step after next
        const size_t bytes = (size_t)s->n * s->vb;
        for (auto &x : s->sh) {
            if (cudaSetDevice(x.dev) != cudaSuccess) { err = CSR5B200_CUDA_ERROR; break; }
            cudaError_t e = cudaSuccess;
            if (!x.x) e = cudaMalloc(&x.x, bytes ? bytes : 1);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(x.x, static_cast<char *>(x.ybuf) + (size_t)s->last * s->stride * s->vb, bytes,
                                    cudaMemcpyDeviceToDevice, x.stream);
            if (e != cudaSuccess) { err = cu(s, e); break; }
            if ((err = csr5b200_set_x(x.h, x.x))) break;
        }
    }
    cudaSetDevice(dev);
    return err;
}

int csr5b200_sharded_synchronize(csr5b200_sharded_t s)
{
    if (!s) return CSR5B200_INVALID_ARGUMENT;
    int dev = 0;
    cudaGetDevice(&dev);
    int err = CSR5B200_SUCCESS;
    for (auto &x : s->sh) {
        if (!x.h) continue;
        cudaSetDevice(x.dev);
        const int e = csr5b200_exchange_status(x.h);   // synchronises the shard's stream
        if (e && !err) err = e;
    }
    cudaSetDevice(dev);
    return err;
}

int csr5b200_sharded_get_y(csr5b200_sharded_t s, int shard, void **y_dev)
{
    if (!s || !y_dev || shard < 0 || shard >= s->G || !s->have_matrix) return CSR5B200_INVALID_ARGUMENT;
    *y_dev = static_cast<char *>(s->sh[shard].ybuf) + (size_t)s->last * s->stride * s->vb;
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_copy_y_to_host(csr5b200_sharded_t s, int shard, void *y_host)
{
    if (!s || !y_host || shard < 0 || shard >= s->G || !s->have_matrix) return CSR5B200_INVALID_ARGUMENT;
    int err = csr5b200_sharded_synchronize(s);
    if (err) return err;
    int dev = 0;
    cudaGetDevice(&dev);
    void *y = nullptr;
    csr5b200_sharded_get_y(s, shard, &y);
    cudaSetDevice(s->sh[shard].dev);
    err = cu(s, cudaMemcpy(y_host, y, (size_t)s->m * s->vb, cudaMemcpyDeviceToHost));
    cudaSetDevice(dev);
    return err;
}

int csr5b200_sharded_get_bounds(csr5b200_sharded_t s, long long *bounds)
{
    if (!s || !bounds || !s->have_matrix) return CSR5B200_INVALID_ARGUMENT;
    for (int g = 0; g <= s->G; g++) bounds[g] = s->bounds[g];
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_get_handle(csr5b200_sharded_t s, int shard, csr5b200_handle_t *h)
{
    if (!s || !h || shard < 0 || shard >= s->G) return CSR5B200_INVALID_ARGUMENT;
    *h = s->sh[shard].h;
    return CSR5B200_SUCCESS;
}

int csr5b200_sharded_destroy(csr5b200_sharded_t s)
{
    if (!s) return CSR5B200_SUCCESS;
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto &x : s->sh) {
        cudaSetDevice(x.dev);
        if (x.stream) cudaStreamSynchronize(x.stream);
    }
    delete s->workers;
    s->workers = nullptr;
    free_matrix(s);
    for (auto &x : s->sh) {
        cudaSetDevice(x.dev);
        if (x.h) csr5b200_free(x.h);
        cudaFree(x.flags);
        if (x.ev_done) cudaEventDestroy(x.ev_done);
        if (x.stream) cudaStreamDestroy(x.stream);
    }
    cudaSetDevice(dev);
    delete s;
    return CSR5B200_SUCCESS;
}

}  // extern "C"
