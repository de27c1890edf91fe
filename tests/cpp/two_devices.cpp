// tests/cpp/two_devices.cpp — one host thread per GPU in ONE process, each bound by gm_init(device), the way
// INTEGRATION.md tells a Go integrator to run one goroutine (locked to an OS thread) per device. Every thread solves
// its own batch and its own waves over its own root, concurrently; results must equal the same work done alone.
// Build: g++ -O2 -std=c++17 -Iinclude tests/cpp/two_devices.cpp -Lgomilp_b200/_build -lgomilp_b200 -lpthread
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "gomilp_b200.h"

struct Batch {
    int m, n, count;
    std::vector<double> c, A, b;
};

static Batch make(int m, int n, int count, unsigned seed) {
    std::mt19937_64 g(seed);
    std::normal_distribution<double> N(0, 1);
    std::uniform_real_distribution<double> U(0, 1);
    Batch B{m, n, count, {}, {}, {}};
    B.c.resize((size_t)count * n); B.A.resize((size_t)count * m * n); B.b.assign((size_t)count * m, 0.0);
    for (int k = 0; k < count; ++k) {
        std::vector<double> x0(n), y0(m);
        for (auto& v : x0) v = U(g) < 0.5 ? U(g) : 0.0;
        for (auto& v : y0) v = N(g);
        double* A = &B.A[(size_t)k * m * n];
        for (int i = 0; i < m * n; ++i) A[i] = N(g);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < n; ++j) B.b[(size_t)k * m + i] += A[i * n + j] * x0[j];
        for (int j = 0; j < n; ++j) {
            double s = U(g);
            for (int i = 0; i < m; ++i) s += A[i * n + j] * y0[i];
            B.c[(size_t)k * n + j] = s;
        }
    }
    return B;
}

struct Out {
    std::vector<int32_t> status;
    std::vector<double> z, x;
    int rc = 0;
};

static Out solve(const Batch& B) {
    Out o;
    o.status.assign(B.count, -1); o.z.assign(B.count, 0); o.x.assign((size_t)B.count * B.n, 0);
    o.rc = gm_simplex_batch(B.count, B.c.data(), B.A.data(), B.b.data(), B.m, B.n, 0.0, o.status.data(), o.z.data(),
                            o.x.data(), nullptr, nullptr);
    return o;
}

int main() {
    const int ndev = gm_device_count();
    if (ndev < 2) { std::printf("SKIP: %d device(s)\n", ndev); return 77; }
    const int T = ndev < 4 ? ndev : 4;
    std::vector<Batch> batches;
    for (int d = 0; d < T; ++d) batches.push_back(make(24 + 20 * d, 60 + 45 * d, 96, 100 + d));
    // reference: every batch alone on device 0
    gm_init(0);
    std::vector<Out> want;
    for (int d = 0; d < T; ++d) want.push_back(solve(batches[d]));
    std::vector<Out> got(T);
    std::vector<int> bad(T, 0);
    std::vector<std::thread> th;
    for (int d = 0; d < T; ++d)
        th.emplace_back([&, d]() {
            if (gm_init(d) != GM_OK) { bad[d] = 1; return; }
            for (int rep = 0; rep < 3; ++rep) {
                got[d] = solve(batches[d]);
                // a root uploaded by this thread lives on this thread's device; a depth-0 wave must equal LP 0
                gm_root_t root = 0;
                const Batch& B = batches[d];
                if (gm_upload_root(B.c.data(), B.A.data(), B.n, B.b.data(), B.m, B.n, &root) != GM_OK) { bad[d] = 2; return; }
                int32_t st = -1; double z = 0; std::vector<double> x(B.n);
                if (gm_solve_wave(root, 1, 0, nullptr, nullptr, nullptr, &st, &z, x.data(), nullptr, nullptr) != GM_OK) bad[d] = 3;
                if (st != got[d].status[0] || z != got[d].z[0]) bad[d] = 4;
                gm_free_root(root);
            }
        });
    for (auto& t : th) t.join();
    int fails = 0;
    for (int d = 0; d < T; ++d) {
        bool same = got[d].rc == GM_OK && want[d].rc == GM_OK && got[d].status == want[d].status && got[d].z == want[d].z &&
                    got[d].x == want[d].x && !bad[d];
        std::printf("thread %d on device %d: %s (flag %d)\n", d, d, same ? "identical to the solo run" : "MISMATCH", bad[d]);
        fails += !same;
    }
    gm_shutdown();
    return fails ? 1 : 0;
}
