// tests/native/collect_check.cpp -- (checker only) C entry points around the host side of the resident kernel's hand-over
// (doppler_b200/csrc/collect.cpp), so that the CPU suite can drive it: units {word, flag} in, result words out.
#include "../../doppler_b200/csrc/collect.h"

extern "C" size_t hostcheck_collect(const uint32_t* units, uint32_t seq, unsigned char* out, size_t from, size_t n)
{
    return dcollect::collect(units, seq, out, from, n);
}
extern "C" size_t hostcheck_collect_scalar(const uint32_t* units, uint32_t seq, unsigned char* out, size_t from, size_t n)
{
    return dcollect::collect_scalar(units, seq, out, from, n);
}

// ---- the whole hand-over against a software device -------------------------------------------------------------------------
// A second thread plays the resident kernel's part of the protocol (mixer_kernels.cuh: mix_resident_kernel, rt_verdict) with the
// product's host code on the other side (dcollect::post / dcollect::collect).  It looks at the four request sectors; a device
// reads a sector as one coherent snapshot, which is modelled here by reading the sector's tag first and its payload words
// afterwards (in a random order, with pauses): on x86 the payload a reader sees after tag n is at least as new as request n's.
// It takes a request when all four tags show the same new number, "mixes" (result word i = input word i * 2654435761 + request
// number) and writes the result units in a random order.  Returns 0 when every request was taken whole, exactly once and in
// order, and every collected result was right; a negative code says what went wrong.
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {
constexpr int kSectors = 4, kPayload = 26;   // as in mixer_kernels.cuh: kRtSectors, kRtPayloadWords
struct Rng {
    uint64_t s;
    uint32_t next()
    {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        return (uint32_t)(s >> 33);
    }
};
}   // namespace

extern "C" int hostcheck_handover(int requests, int max_words, uint32_t seed, uint32_t first_seq)
{
    alignas(64) static volatile uint32_t mailbox[kSectors * 8];
    std::vector<uint32_t> units(2 * (size_t)max_words, 0), input((size_t)max_words), out((size_t)max_words);
    std::vector<uint32_t> in_shared((size_t)max_words);   // the "staging buffer" the device reads
    for (int i = 0; i < kSectors * 8; i++) mailbox[i] = 0;
    std::atomic<int> device_error{0};
    std::atomic<bool> stop{false};
    std::atomic<uint32_t> served_count{0};
    uint32_t last0 = first_seq - 1u;
    if (last0 == 0) last0 = 0xffffffffu;   // (numbers skip 0; the mailbox starts at 0)

    std::thread device([&] {
        Rng rng{seed ^ 0x9e3779b97f4a7c15ull};
        uint32_t last = last0;
        while (!stop.load(std::memory_order_relaxed)) {
            // one look: the four sectors in a random order, each tag first
            uint32_t tags[kSectors], words[kSectors * 7];
            int sorder[kSectors] = {0, 1, 2, 3};
            for (int i = kSectors - 1; i > 0; i--) std::swap(sorder[i], sorder[rng.next() % (i + 1)]);
            for (int si = 0; si < kSectors; si++) {
                const int t = sorder[si];
                tags[t] = mailbox[8 * t + 7];
                std::atomic_thread_fence(std::memory_order_acquire);
                int order[7] = {0, 1, 2, 3, 4, 5, 6};
                for (int i = 6; i > 0; i--) std::swap(order[i], order[rng.next() % (i + 1)]);
                for (int i = 0; i < 7; i++) {
                    words[7 * t + order[i]] = mailbox[8 * t + order[i]];
                    if (rng.next() % 8 == 0) std::this_thread::yield();
                }
            }
            const uint32_t seq = tags[0];
            if (tags[1] != seq || tags[2] != seq || tags[3] != seq || seq == last) continue;
            // a whole request: words[0] = result words, words[1] = request number as the host believes it, words[2..] = echo
            const uint32_t n = words[0];
            if (words[1] != seq) { device_error = -11; return; }                       // a torn request got through
            for (int i = 2; i < kPayload; i++)
                if (words[i] != seq * 31u + (uint32_t)i) { device_error = -12; return; }
            uint32_t expect = last + 1u;
            if (expect == 0) expect = 1;
            if (seq != expect) { device_error = -13; return; }                          // skipped or repeated
            last = seq;
            std::vector<uint32_t> idx(n);
            for (uint32_t i = 0; i < n; i++) idx[i] = i;
            for (uint32_t i = n; i > 1; i--) std::swap(idx[i - 1], idx[rng.next() % i]);
            for (uint32_t j = 0; j < n; j++) {
                const uint32_t i = idx[j];
                const uint64_t unit = ((uint64_t)seq << 32) | (uint32_t)(in_shared[i] * 2654435761u + seq);
                reinterpret_cast<std::atomic<uint64_t>*>(units.data())[i].store(unit, std::memory_order_relaxed);   // one 8-byte store
            }
            served_count.fetch_add(1, std::memory_order_relaxed);
        }
    });

    Rng rng{seed};
    uint32_t seq = first_seq - 1u;
    int rc = 0;
    for (int r = 0; r < requests && rc == 0; r++) {
        if (++seq == 0) {
            std::fill(units.begin(), units.end(), 0u);   // as rt_post does when the numbers wrap
            ++seq;
        }
        const uint32_t n = 1 + rng.next() % (uint32_t)max_words;
        for (uint32_t i = 0; i < n; i++) input[i] = rng.next();
        memcpy(in_shared.data(), input.data(), n * 4);
        uint32_t payload[kPayload];
        payload[0] = n;
        payload[1] = seq;
        for (int i = 2; i < kPayload; i++) payload[i] = seq * 31u + (uint32_t)i;
        dcollect::post(mailbox, kSectors, seq, payload, kPayload);
        size_t got = 0;
        for (uint64_t spins = 0; got < n; spins++) {
            got = dcollect::collect(units.data(), seq, reinterpret_cast<unsigned char*>(out.data()), got, n);
            if (device_error.load()) { rc = device_error.load(); break; }
            if (spins > 2000000000ull) { rc = -1; break; }   // hung
            if ((spins & 1023) == 1023) std::this_thread::yield();
        }
        for (uint32_t i = 0; rc == 0 && i < n; i++)
            if (out[i] != input[i] * 2654435761u + seq) rc = -2;   // a stale or foreign word was handed out
    }
    stop = true;
    device.join();
    if (rc == 0 && device_error.load()) rc = device_error.load();
    if (rc == 0 && served_count.load() != (uint32_t)requests) rc = -3;
    return rc;
}

// the request layout as posted (dcollect::post): 32 words for four sectors
extern "C" void hostcheck_post(uint32_t* mailbox, int sectors, uint32_t seq, const uint32_t* payload, int payload_words)
{
    dcollect::post(mailbox, sectors, seq, payload, payload_words);
}
