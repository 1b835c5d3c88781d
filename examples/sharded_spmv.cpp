// sharded_spmv.cpp -- the multi-GPU SpMV from C++, no Python, no torch: include/csr5_b200_sharded.h.
//
//   g++ -O2 -std=c++17 -Iinclude examples/sharded_spmv.cpp -Lbenchmark_spmv_using_csr5_b200 -lcsr5_b200
//       -Wl,-rpath,$PWD/benchmark_spmv_using_csr5_b200 -o sharded_spmv
//   ./sharded_spmv [shards] [rows] [steps] [transport 0..5] [same-device 0|1]
//
// Builds the banded test matrix of BASELINE.json configs[1] (16 nnz/row, wrap-around) on the host -- the CSR the
// reference's loader would hand to inputCSR (CSR5_cuda/main.cu:211-306) -- splits it over `shards` GPUs (or `shards`
// shards on GPU 0 with same-device = 1), runs `steps` steps, checks every device's gathered y against the scalar CSR
// loop of the reference (main.cu:336-350) and prints the step time.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "csr5_b200_sharded.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        const int err__ = (call);                                                                     \
        if (err__) { std::fprintf(stderr, "%s -> %d (%s)\n", #call, err__, csr5b200_error_string(err__)); return 1; } \
    } while (0)

int main(int argc, char **argv)
{
    const int shards = argc > 1 ? std::atoi(argv[1]) : 2;
    const int m = argc > 2 ? std::atoi(argv[2]) : 1000000;
    const int steps = argc > 3 ? std::atoi(argv[3]) : 20;
    const int transport = argc > 4 ? std::atoi(argv[4]) : CSR5B200_TRANSPORT_AUTO;
    const int same_device = argc > 5 ? std::atoi(argv[5]) : 0;
    const int per_row = 16;
    const int nnz = m * per_row;

    std::vector<int> row_ptr(m + 1), col((size_t)nnz);
    std::vector<double> val((size_t)nnz), x(m), y_ref(m), y(m);
    unsigned long long s = 42;
    auto rnd = [&]() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (double)((s >> 33) % 10); };
    for (int i = 0; i <= m; i++) row_ptr[i] = i * per_row;
    for (int i = 0; i < m; i++)
        for (int k = 0; k < per_row; k++) {
            col[(size_t)i * per_row + k] = (int)(((long long)i - per_row / 2 + k + m) % m);
            val[(size_t)i * per_row + k] = rnd();   // integer-valued, as the reference's rand() % 10 (main.cu:314-326)
        }
    for (int i = 0; i < m; i++) x[i] = rnd();
    for (int i = 0; i < m; i++) {   // main.cu:343-349
        double sum = 0;
        for (int j = row_ptr[i]; j < row_ptr[i + 1]; j++) sum += x[col[j]] * val[j];
        y_ref[i] = sum;
    }

    std::vector<int> devices(shards);
    for (int g = 0; g < shards; g++) devices[g] = same_device ? 0 : g;
    csr5b200_sharded_t A = nullptr;
    CHECK(csr5b200_sharded_create(shards, devices.data(), 8, &A));
    CHECK(csr5b200_sharded_input_csr_host(A, m, m, nnz, row_ptr.data(), col.data(), val.data()));
    CHECK(csr5b200_sharded_set_exchange(A, transport, 0, 0, CSR5B200_BARRIER_AUTO, 0));
    CHECK(csr5b200_sharded_set_x_host(A, x.data()));
    CHECK(csr5b200_sharded_as_csr5(A));
    for (int w = 0; w < 3; w++) CHECK(csr5b200_sharded_spmv(A, 1.0, 0.0));
    CHECK(csr5b200_sharded_synchronize(A));
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < steps; k++) CHECK(csr5b200_sharded_spmv(A, 1.0, 0.0));
    CHECK(csr5b200_sharded_synchronize(A));
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / steps;

    long long bad = 0;
    for (int g = 0; g < shards; g++) {
        CHECK(csr5b200_sharded_copy_y_to_host(A, g, y.data()));
        for (int i = 0; i < m; i++) bad += y[i] != y_ref[i];
    }
    std::vector<long long> bounds(shards + 1);
    CHECK(csr5b200_sharded_get_bounds(A, bounds.data()));
    std::printf("%s: %d shards%s, %d x %d, %d nnz; rows per shard:", csr5b200_version(), shards,
                same_device ? " on device 0" : "", m, m, nnz);
    for (int g = 0; g < shards; g++) std::printf(" %lld", bounds[g + 1] - bounds[g]);
    std::printf("\nstep (SpMV + y on every device) = %.4f ms, %.1f GFlops (host-timed over %d steps)\n", ms,
                2.0 * nnz / (ms * 1e6), steps);
    std::printf("Check... %s\n", bad == 0 ? "PASS!" : "NO PASS!");
    CHECK(csr5b200_sharded_destroy(A));
    return bad == 0 ? 0 : 3;
}
