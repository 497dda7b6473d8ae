// doppler_b200/csrc/collect.cpp -- see collect.h.  Plain C++ (AVX2 where the CPU has it), no CUDA.
#include "collect.h"

#include <string.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>

namespace dcollect {

void post(volatile uint32_t* mailbox, int sectors, uint32_t seq, const uint32_t* payload, int payload_words)
{
    constexpr int kW = 7;   // payload words per sector
    for (int t = 0; t < sectors; t++) {
        volatile uint32_t* sector = mailbox + 8 * t;
        for (int i = 0; i < kW; i++) {
            const int w = t * kW + i;
            sector[i] = w < payload_words ? payload[w] : 0u;
        }
        std::atomic_thread_fence(std::memory_order_release);   // (x86 keeps stores in order; this keeps the compiler from moving them)
        sector[kW] = seq;
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);   // out of the store buffer now, not when the caller's spin loop lets it
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) static size_t collect_avx2(const uint32_t* units, uint32_t seq, unsigned char* out, size_t from, size_t n)
{
    const __m256i vseq = _mm256_set1_epi32((int)seq);
    const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 1, 3, 5, 7);
    size_t i = from;
    for (; i + 8 <= n; i += 8) {   // eight units = one 64-byte line of the device's writes
        // (every aligned 8-byte unit inside a vector load is read whole, which is all the scheme needs: word and flag together)
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(units + 2 * i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(units + 2 * i + 8));
        const __m256i pa = _mm256_permutevar8x32_epi32(a, idx), pb = _mm256_permutevar8x32_epi32(b, idx);   // words | flags
        const __m256i flags = _mm256_permute2x128_si256(pa, pb, 0x31);
        if (_mm256_movemask_epi8(_mm256_cmpeq_epi32(flags, vseq)) != -1) break;
        _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + 4 * i), _mm256_permute2x128_si256(pa, pb, 0x20));
    }
    return i;
}
#endif

size_t collect_scalar(const uint32_t* units, uint32_t seq, unsigned char* out, size_t from, size_t n)
{
    size_t i = from;
    for (; i < n; i++) {
        const uint64_t u = *reinterpret_cast<const volatile uint64_t*>(units + 2 * i);   // one aligned 8-byte load
        if ((uint32_t)(u >> 32) != seq) break;
        const uint32_t w = (uint32_t)u;
        memcpy(out + 4 * i, &w, 4);
    }
    return i;
}

size_t collect(const uint32_t* units, uint32_t seq, unsigned char* out, size_t from, size_t n)
{
    asm volatile("" ::: "memory");   // the units are written by the device: read them again on every call
    size_t i = from;
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) {
        i = collect_avx2(units, seq, out, i, n);
        if (i + 8 <= n) return i;   // stopped at a line that is not there yet
    }
#endif
    return collect_scalar(units, seq, out, i, n);
}

}   // namespace dcollect
