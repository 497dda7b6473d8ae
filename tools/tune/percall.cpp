// tools/tune/percall.cpp -- latency of the host-buffer C-ABI calls at the reference's own granularity
// (one 8192-byte block per call, src/main.rs:49,70) and at the batched sizes INTEGRATION.md recommends,
// with the resident kernel (default), with a zero-copy launch per block, and with the staged pipeline; pageable and pinned caller buffers.
// Build: make -C tools/tune percall      Run (GPU box): tools/tune/percall
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
    const bool quick = getenv("PERCALL_QUICK") != nullptr;   // bench.py: the 8192-byte block only, resident kernel vs launch per block
    const float shift = getenv("PERCALL_SHIFT") ? (float)atof(getenv("PERCALL_SHIFT")) : 5000.0f;   // (7321.7: a period beyond any table)
    const size_t quick_n = getenv("PERCALL_N") ? (size_t)atol(getenv("PERCALL_N")) : 2048;   // (2047: the plan differs from block to block)
    const std::vector<size_t> sizes = quick ? std::vector<size_t>{quick_n} : std::vector<size_t>{1024, 2048, 16384, 262144, 4194304, 33554432};   // complex samples per call
    for (int mode = 2; mode >= (quick ? 1 : 0); mode--) {   // 2: resident kernel (blocks up to 32 KiB), 1: one zero-copy launch per block, 0: staged pipeline
        const int tiny = mode >= 1;
        doppler_b200_tune(ctx, DOPPLER_B200_TUNE_TINY_HOST_BYTES, tiny ? (128u << 10) : 0);
        doppler_b200_tune(ctx, DOPPLER_B200_TUNE_RESIDENT_IDLE_US, mode == 2 ? 20000 : 0);
        for (int pinned = 0; pinned < (quick ? 1 : 2); pinned++)
            for (size_t n : sizes) {
                if (pinned && n > 262144) continue;
                if (mode == 1 && n > 8192) continue;   // (identical to mode 2 above 32 KiB)
                const size_t bytes = n * 4;   // i16 IQ in and out
                void *in, *out;
                if (pinned) {
                    in = doppler_b200_host_alloc(bytes);
                    out = doppler_b200_host_alloc(bytes);
                } else {
                    in = malloc(bytes);
                    out = malloc(bytes);
                }
                memset(in, 1, bytes);
                uint32_t sn = 0;
                size_t got = 0;
                const int iters = n <= 16384 ? 5000 : n <= 262144 ? 1000 : n <= 4194304 ? 100 : 20;
                for (int i = 0; i < 50; i++) doppler_b200_mix(ctx, in, bytes, DOPPLER_B200_I16, DOPPLER_B200_I16, shift, 1024000, &sn, out, bytes, &got);
                auto t0 = now();
                for (int i = 0; i < iters; i++)
                    if (doppler_b200_mix(ctx, in, bytes, DOPPLER_B200_I16, DOPPLER_B200_I16, shift, 1024000, &sn, out, bytes, &got) != 0) return 3;
                const double us = std::chrono::duration<double, std::micro>(now() - t0).count() / iters;
                printf("{\"call\": \"doppler_b200_mix i16->i16\", \"path\": \"%s\", \"caller_buffers\": \"%s\", \"samples_per_call\": %zu, "
                       "\"us_per_call\": %.2f, \"msps\": %.1f}\n",
                       mode == 2 ? "resident kernel / zero-copy" : mode == 1 ? "zero-copy launch per block" : "staged pipeline", pinned ? "pinned" : "pageable", n, us, n / us);
                fflush(stdout);
                if (pinned) {
                    doppler_b200_host_free(in);
                    doppler_b200_host_free(out);
                } else {
                    free(in);
                    free(out);
                }
            }
    }
    // Paced: a realtime stream delivers one block every few milliseconds, not back to back.  Calls 30-60 us apart (the
    // resident kernel stays; a random phase against its looks), the call's own duration timed -- mean and median.
    for (int mode = 2; mode >= 1; mode--) {
        doppler_b200_tune(ctx, DOPPLER_B200_TUNE_TINY_HOST_BYTES, 128u << 10);
        doppler_b200_tune(ctx, DOPPLER_B200_TUNE_RESIDENT_IDLE_US, mode == 2 ? 20000 : 0);
        const size_t n = quick ? quick_n : 2048, bytes = n * 4;
        void *in = malloc(bytes), *out = malloc(bytes);
        memset(in, 1, bytes);
        uint32_t sn = 0, lcg = 12345;
        size_t got = 0;
        const int iters = quick ? 3000 : 10000;
        std::vector<double> us(iters);
        for (int i = -50; i < iters; i++) {
            lcg = lcg * 1664525u + 1013904223u;
            const double gap_us = 30.0 + (lcg >> 8) * (30.0 / (1u << 24));
            const auto g0 = now();
            while (std::chrono::duration<double, std::micro>(now() - g0).count() < gap_us) {
            }
            const auto t0 = now();
            if (doppler_b200_mix(ctx, in, bytes, DOPPLER_B200_I16, DOPPLER_B200_I16, shift, 1024000, &sn, out, bytes, &got) != 0) return 3;
            if (i >= 0) us[i] = std::chrono::duration<double, std::micro>(now() - t0).count();
        }
        double mean = 0;
        for (double u : us) mean += u / iters;
        std::sort(us.begin(), us.end());
        printf("{\"call\": \"doppler_b200_mix i16->i16\", \"path\": \"%s\", \"caller_buffers\": \"pageable\", \"samples_per_call\": %zu, "
               "\"paced\": \"30-60 us between calls\", \"us_per_call\": %.2f, \"median_us\": %.2f, \"p99_us\": %.2f}\n",
               mode == 2 ? "resident kernel / zero-copy" : "zero-copy launch per block", n, mean, us[iters / 2], us[iters * 99 / 100]);
        fflush(stdout);
        free(in);
        free(out);
    }
    doppler_b200_destroy(ctx);
    return 0;
}
