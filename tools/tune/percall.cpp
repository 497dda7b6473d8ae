// tools/tune/percall.cpp -- latency of the host-buffer C-ABI calls at the reference's own granularity
// (one 8192-byte block per call, src/main.rs:49,70) and at the batched sizes INTEGRATION.md recommends.
// Build: make -C tools/tune percall      Run (GPU box): tools/tune/percall
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/doppler_b200.h"

int main()
{
    doppler_b200_ctx* ctx = nullptr;
    if (doppler_b200_create(0, &ctx) != 0) {
        fprintf(stderr, "create failed: %s\n", doppler_b200_last_error(nullptr));
        return 2;
    }
    auto now = [] { return std::chrono::steady_clock::now(); };
    const size_t sizes[] = {1024, 16384, 262144, 4194304, 33554432};   // complex samples per call
    for (size_t n : sizes) {
        std::vector<float> in(2 * n, 0.25f), out(2 * n);
        uint32_t sn = 0;
        const int iters = n <= 16384 ? 4000 : n <= 262144 ? 1000 : n <= 4194304 ? 100 : 20;
        for (int i = 0; i < 20; i++) doppler_b200_shift_frequency(ctx, in.data(), n, &sn, 815000.0f, 2400000, out.data());
        auto t0 = now();
        for (int i = 0; i < iters; i++)
            if (doppler_b200_shift_frequency(ctx, in.data(), n, &sn, 815000.0f, 2400000, out.data()) != 0) return 3;
        const double us = std::chrono::duration<double, std::micro>(now() - t0).count() / iters;
        printf("{\"call\": \"doppler_b200_shift_frequency (pageable host buffers)\", \"samples_per_call\": %zu, \"us_per_call\": %.1f, "
               "\"msps\": %.1f}\n", n, us, n / us);
    }
    doppler_b200_destroy(ctx);
    return 0;
}
