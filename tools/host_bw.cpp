// host_bw.cpp — what the host of a GPU box can stream: threads reading one large buffer (the raw
// depth frames the run-length rewrite of dh_predict_batch has to scan).  g++ -O2 -pthread.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <thread>
#include <vector>
int main() {
    const size_t N = 1024ull * 614400 / 8;  // 629 MB: the 1024 frames of a bench step
    std::vector<uint64_t> a(N, 1);
    for (unsigned t : {1u, 2u, 4u, 8u, 12u, 16u, 24u, 32u}) {
        if (t > std::thread::hardware_concurrency()) break;
        double best = 1e9;
        for (int r = 0; r < 3; ++r) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            std::vector<uint64_t> s(t);
            for (unsigned k = 0; k < t; ++k)
                th.emplace_back([&, k] {
                    uint64_t x = 0;
                    for (size_t i = k * N / t; i < (k + 1) * N / t; ++i) x |= a[i];
                    s[k] = x;
                });
            for (auto& x : th) x.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (dt < best) best = dt;
        }
        printf("{\"threads\": %u, \"read_GBps\": %.1f}\n", t, N * 8 / best / 1e9);
    }
    return 0;
}
