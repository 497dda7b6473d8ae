// multi.cpp -- time-sliced multi-GPU mixing behind the C ABI (BASELINE north_star:
// "partition across the 8 GPUs of one box by contiguous time-slice with analytically-carried
// starting phase").
//
// The path shards with NO exchange step.  One stream is cut into contiguous slices on whole pump
// blocks (/root/reference/src/main.rs:49: 8192 bytes of input, so the per-block shift schedule of
// track mode, main.rs:177, stays aligned); the only cross-slice state is the reference's
// `samplenum` (main.rs:60) at the first sample of each slice, which the host planner (plan.h)
// carries analytically -- O(period search), not O(samples).  Slice d is mixed by device d through
// that device's own context (doppler_b200.cu: streams, 3-slot host pipeline, table arena), driven by
// one persistent host thread per device so that planning, staging copies and launches of the slices
// run side by side.  No collective, no peer access: every device reads its slice and writes its
// slice.  Host-only translation unit: it uses the public C ABI and the planner, nothing else.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/doppler_b200.h"
#include "plan.h"

namespace {

inline size_t bps_of(int t) { return t == DOPPLER_B200_I16 ? 4 : 8; }
inline bool valid_type(int t) { return t == DOPPLER_B200_I16 || t == DOPPLER_B200_F32; }

// [begin, end) of slice `index`: whole blocks, the remainder blocks to the lowest slices, the ragged
// tail (a short last block) to the last slice.
void bounds(uint64_t total, uint32_t nslices, uint32_t index, uint64_t block_samples, uint64_t* begin, uint64_t* end)
{
    const uint64_t nblocks = total / block_samples, per = nblocks / nslices, extra = nblocks % nslices;
    const uint64_t b0 = index * per + (index < extra ? index : extra);
    const uint64_t b1 = b0 + per + (index < extra ? 1 : 0);
    *begin = b0 * block_samples;
    *end = index + 1 == nslices ? total : b1 * block_samples;
}

// runs of the stream samples [begin, begin + count) under a per-block schedule (nblocks == 1: the one
// shift covers the whole stream, const mode)
std::vector<dplan::Run> runs_of(const float* shifts, size_t nblocks, uint64_t block_samples, uint32_t samplerate, uint64_t begin,
                                uint64_t count)
{
    if (nblocks == 1) return {dplan::Run{count, dplan::ratio(shifts[0], samplerate)}};
    const size_t b0 = (size_t)(begin / block_samples);
    if (b0 >= nblocks) return {};
    return dplan::runs_from_blocks(shifts + b0, nblocks - b0, block_samples, samplerate, count);
}

}  // namespace

struct doppler_b200_multi {
    std::vector<int> devices;
    std::vector<doppler_b200_ctx*> ctx;
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    uint64_t generation = 0;
    int pending = 0;
    bool quit = false;
    std::function<int(int)> job;
    std::vector<int> rc;
    std::string err;
    dplan::Planner planner;   // seeds of the slices (caller's thread)

    // runs job(d) on every device's thread, returns the first non-zero status
    int run(std::function<int(int)> fn)
    {
        std::unique_lock<std::mutex> g(m);
        job = std::move(fn);
        pending = (int)devices.size();
        generation++;
        cv_work.notify_all();
        cv_done.wait(g, [&] { return pending == 0; });
        for (size_t d = 0; d < rc.size(); d++)
            if (rc[d]) {
                char buf[600];
                snprintf(buf, sizeof buf, "device %d: %s", devices[d], ctx[d] ? doppler_b200_last_error(ctx[d]) : "no context");
                err = buf;
                return rc[d];
            }
        return DOPPLER_B200_OK;
    }

    void worker(int d)
    {
        uint64_t seen = 0;
        for (;;) {
            std::function<int(int)> fn;
            {
                std::unique_lock<std::mutex> g(m);
                cv_work.wait(g, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
                fn = job;
            }
            const int r = fn(d);
            {
                std::lock_guard<std::mutex> g(m);
                rc[d] = r;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
};

namespace {

int mfail(doppler_b200_multi* m, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m) m->err = buf;
    return code;
}

struct SlicePlan {
    std::vector<uint64_t> begin;   // nslices + 1 entries
    std::vector<uint32_t> seed;    // nslices + 1 entries: seed[i] = samplenum at begin[i]; the last = state after the stream
};

// Slices of `total` samples on block boundaries and the samplenum carried to each of them.
SlicePlan plan_slices(dplan::Planner& pl, uint32_t samplenum, const float* shifts, size_t nblocks, uint64_t block_samples,
                      uint32_t samplerate, uint64_t total, uint32_t nslices)
{
    SlicePlan sp;
    sp.begin.resize(nslices + 1);
    sp.seed.resize(nslices + 1);
    uint32_t sn = samplenum;
    for (uint32_t i = 0; i < nslices; i++) {
        uint64_t b, e;
        bounds(total, nslices, i, block_samples, &b, &e);
        sp.begin[i] = b;
        sp.seed[i] = sn;
        sn = pl.advance(runs_of(shifts, nblocks, block_samples, samplerate, b, e - b), sn);
    }
    sp.begin[nslices] = total;
    sp.seed[nslices] = sn;
    return sp;
}

int check_stream(doppler_b200_multi* m, size_t in_len, int intype, int outtype, const float* shifts, size_t nblocks, size_t block_bytes,
                 uint32_t* samplenum, uint64_t* nsamples, uint64_t* block_samples)
{
    if (!m) return DOPPLER_B200_EINVAL;
    if (!valid_type(intype) || !valid_type(outtype)) return mfail(m, DOPPLER_B200_EINVAL, "unknown IQ data type");
    if (!samplenum) return mfail(m, DOPPLER_B200_EINVAL, "samplenum is NULL");
    if (!shifts || nblocks == 0) return mfail(m, DOPPLER_B200_EINVAL, "no shift schedule");
    const size_t ibps = bps_of(intype);
    if (in_len % ibps != 0)
        return mfail(m, DOPPLER_B200_EALIGN, "input length %zu is not a multiple of %zu (dsp.rs assert)", in_len, ibps);
    if (block_bytes == 0 || block_bytes % ibps != 0)
        return mfail(m, DOPPLER_B200_EINVAL, "block_bytes %zu is not a whole number of samples", block_bytes);
    if (nblocks > 1 && (in_len + block_bytes - 1) / block_bytes > nblocks)
        return mfail(m, DOPPLER_B200_EINVAL, "shift schedule has %zu blocks, input needs %zu", nblocks,
                     (in_len + block_bytes - 1) / block_bytes);
    *nsamples = in_len / ibps;
    *block_samples = block_bytes / ibps;
    return DOPPLER_B200_OK;
}

int mix_multi_host(doppler_b200_multi* m, const void* in, size_t in_len, int intype, int outtype, const float* shifts, size_t nblocks,
                   size_t block_bytes, uint32_t samplerate, uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len)
{
    uint64_t n = 0, bs = 0;
    int rc = check_stream(m, in_len, intype, outtype, shifts, nblocks, block_bytes, samplenum, &n, &bs);
    if (rc) return rc;
    if (out_len) *out_len = 0;
    const size_t ibps = bps_of(intype), obps = bps_of(outtype);
    if (n && (!in || !out)) return mfail(m, DOPPLER_B200_EINVAL, "NULL buffer");
    if (n * obps > out_cap) return mfail(m, DOPPLER_B200_ECAP, "output capacity %zu < %llu bytes needed", out_cap, (unsigned long long)(n * obps));
    if (n == 0) return DOPPLER_B200_OK;
    const uint32_t ns = (uint32_t)m->devices.size();
    const SlicePlan sp = plan_slices(m->planner, *samplenum, shifts, nblocks, bs, samplerate, n, ns);
    rc = m->run([&](int d) -> int {
        const uint64_t b = sp.begin[d], e = sp.begin[d + 1];
        if (e == b) return DOPPLER_B200_OK;
        uint32_t sn = sp.seed[d];
        const char* src = static_cast<const char*>(in) + b * ibps;
        char* dst = static_cast<char*>(out) + b * obps;
        if (nblocks == 1)
            return doppler_b200_mix(m->ctx[d], src, (e - b) * ibps, intype, outtype, shifts[0], samplerate, &sn, dst, (e - b) * obps, nullptr);
        const size_t b0 = (size_t)(b / bs);
        return doppler_b200_mix_blocks(m->ctx[d], src, (e - b) * ibps, intype, outtype, shifts + b0, nblocks - b0, block_bytes, samplerate, &sn,
                                       dst, (e - b) * obps, nullptr);
    });
    if (rc) return rc;
    *samplenum = sp.seed[ns];
    if (out_len) *out_len = n * obps;
    return DOPPLER_B200_OK;
}

int mix_multi_dev(doppler_b200_multi* m, const void* const* d_in, const size_t* in_len, int intype, int outtype, const float* shifts,
                  size_t nblocks, size_t block_bytes, uint32_t samplerate, uint32_t* samplenum, void* const* d_out, const size_t* out_cap)
{
    if (!m) return DOPPLER_B200_EINVAL;
    if (!d_in || !in_len || !d_out || !out_cap) return mfail(m, DOPPLER_B200_EINVAL, "NULL slice arrays");
    const size_t ns = m->devices.size();
    size_t total_len = 0;
    for (size_t d = 0; d < ns; d++) total_len += in_len[d];
    uint64_t n = 0, bs = 0;
    int rc = check_stream(m, total_len, intype, outtype, shifts, nblocks, block_bytes, samplenum, &n, &bs);
    if (rc) return rc;
    // the caller chose the slices: every slice but the last non-empty one must be whole blocks
    std::vector<uint64_t> begin(ns + 1, 0);
    std::vector<uint32_t> seed(ns + 1, *samplenum);
    const size_t ibps = bps_of(intype);
    size_t last = 0;
    for (size_t d = 0; d < ns; d++)
        if (in_len[d]) last = d;
    for (size_t d = 0; d < ns; d++) {
        if (in_len[d] % ibps != 0) return mfail(m, DOPPLER_B200_EALIGN, "slice %zu: length %zu is not a whole number of samples", d, in_len[d]);
        if (nblocks > 1 && d != last && in_len[d] % block_bytes != 0)
            return mfail(m, DOPPLER_B200_EINVAL, "slice %zu: %zu bytes is not a whole number of %zu-byte blocks", d, in_len[d], block_bytes);
        const uint64_t cnt = in_len[d] / ibps;
        begin[d + 1] = begin[d] + cnt;
        seed[d + 1] = m->planner.advance(runs_of(shifts, nblocks, bs, samplerate, begin[d], cnt), seed[d]);
    }
    rc = m->run([&](int d) -> int {
        if (in_len[d] == 0) return DOPPLER_B200_OK;
        uint32_t sn = seed[d];
        if (nblocks == 1)
            return doppler_b200_mix_dev(m->ctx[d], d_in[d], in_len[d], intype, outtype, shifts[0], samplerate, &sn, d_out[d], out_cap[d], nullptr);
        const size_t b0 = (size_t)(begin[d] / bs);
        return doppler_b200_mix_blocks_dev(m->ctx[d], d_in[d], in_len[d], intype, outtype, shifts + b0, nblocks - b0, block_bytes, samplerate,
                                           &sn, d_out[d], out_cap[d], nullptr);
    });
    if (rc) return rc;
    *samplenum = seed[ns];
    return DOPPLER_B200_OK;
}

}  // namespace

extern "C" {

int doppler_b200_slice_bounds(uint64_t total_samples, uint32_t nslices, uint32_t index, uint64_t block_samples, uint64_t* begin,
                              uint64_t* end)
{
    if (nslices == 0 || index >= nslices || block_samples == 0 || !begin || !end) return DOPPLER_B200_EINVAL;
    bounds(total_samples, nslices, index, block_samples, begin, end);
    return DOPPLER_B200_OK;
}

int doppler_b200_slice_seeds(uint32_t samplenum, const float* shift_hz_per_block, size_t nblocks, uint64_t block_samples,
                             uint32_t samplerate, uint64_t total_samples, uint32_t nslices, uint64_t* begins, uint32_t* seeds)
{
    if (!shift_hz_per_block || nblocks == 0 || block_samples == 0 || nslices == 0 || !begins || !seeds) return DOPPLER_B200_EINVAL;
    if (nblocks > 1 && (total_samples + block_samples - 1) / block_samples > nblocks) return DOPPLER_B200_EINVAL;
    dplan::Planner pl;
    const SlicePlan sp = plan_slices(pl, samplenum, shift_hz_per_block, nblocks, block_samples, samplerate, total_samples, nslices);
    for (uint32_t i = 0; i <= nslices; i++) {
        begins[i] = sp.begin[i];
        seeds[i] = sp.seed[i];
    }
    return DOPPLER_B200_OK;
}

int doppler_b200_multi_create(const int* devices, int ndevices, doppler_b200_multi** out)
{
    if (!out || ndevices < 0 || ndevices > 64) return DOPPLER_B200_EINVAL;
    *out = nullptr;
    doppler_b200_multi* m = new (std::nothrow) doppler_b200_multi;
    if (!m) return DOPPLER_B200_ENOMEM;
    if (ndevices == 0) {
        // all visible devices: probe by creating contexts until the ordinal runs out
        for (int d = 0; d < 64; d++) {
            doppler_b200_ctx* c = nullptr;
            if (doppler_b200_create(d, &c) != DOPPLER_B200_OK) break;
            m->devices.push_back(d);
            m->ctx.push_back(c);
        }
        if (m->devices.empty()) {
            delete m;
            return DOPPLER_B200_ENODEV;   // doppler_b200_last_error(NULL) has the reason
        }
    } else {
        for (int i = 0; i < ndevices; i++) {
            const int d = devices ? devices[i] : i;
            doppler_b200_ctx* c = nullptr;
            const int rc = doppler_b200_create(d, &c);
            if (rc != DOPPLER_B200_OK) {
                for (doppler_b200_ctx* x : m->ctx) doppler_b200_destroy(x);
                delete m;
                return rc;
            }
            m->devices.push_back(d);
            m->ctx.push_back(c);
        }
    }
    m->rc.assign(m->devices.size(), 0);
    for (size_t d = 0; d < m->devices.size(); d++) m->threads.emplace_back([m, d] { m->worker((int)d); });
    *out = m;
    return DOPPLER_B200_OK;
}

void doppler_b200_multi_destroy(doppler_b200_multi* m)
{
    if (!m) return;
    {
        std::lock_guard<std::mutex> g(m->m);
        m->quit = true;
        m->cv_work.notify_all();
    }
    for (std::thread& t : m->threads) t.join();
    for (doppler_b200_ctx* c : m->ctx) doppler_b200_destroy(c);
    delete m;
}

int doppler_b200_multi_size(const doppler_b200_multi* m) { return m ? (int)m->devices.size() : 0; }

doppler_b200_ctx* doppler_b200_multi_ctx(doppler_b200_multi* m, int index)
{
    return (m && index >= 0 && (size_t)index < m->ctx.size()) ? m->ctx[index] : nullptr;
}

const char* doppler_b200_multi_last_error(const doppler_b200_multi* m) { return m ? m->err.c_str() : doppler_b200_last_error(nullptr); }

uint64_t doppler_b200_multi_launch_count(const doppler_b200_multi* m)
{
    uint64_t n = 0;
    if (m)
        for (doppler_b200_ctx* c : m->ctx) n += doppler_b200_launch_count(c);
    return n;
}

int doppler_b200_multi_synchronize(doppler_b200_multi* m)
{
    if (!m) return DOPPLER_B200_EINVAL;
    return m->run([&](int d) { return doppler_b200_synchronize(m->ctx[d]); });
}

int doppler_b200_mix_multi(doppler_b200_multi* m, const void* in, size_t in_len, int intype, int outtype, float shift_hz,
                           uint32_t samplerate, uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len)
{
    return mix_multi_host(m, in, in_len, intype, outtype, &shift_hz, 1, DOPPLER_B200_BUFFER_SIZE, samplerate, samplenum, out, out_cap, out_len);
}

int doppler_b200_mix_blocks_multi(doppler_b200_multi* m, const void* in, size_t in_len, int intype, int outtype,
                                  const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                  uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len)
{
    if (m && nblocks == 1 && in_len > block_bytes) return mfail(m, DOPPLER_B200_EINVAL, "shift schedule has 1 block, input needs more");
    return mix_multi_host(m, in, in_len, intype, outtype, shift_hz_per_block, nblocks, block_bytes, samplerate, samplenum, out, out_cap, out_len);
}

int doppler_b200_mix_multi_dev(doppler_b200_multi* m, const void* const* d_in, const size_t* in_len, int intype, int outtype,
                               float shift_hz, uint32_t samplerate, uint32_t* samplenum, void* const* d_out, const size_t* out_cap)
{
    return mix_multi_dev(m, d_in, in_len, intype, outtype, &shift_hz, 1, DOPPLER_B200_BUFFER_SIZE, samplerate, samplenum, d_out, out_cap);
}

int doppler_b200_mix_blocks_multi_dev(doppler_b200_multi* m, const void* const* d_in, const size_t* in_len, int intype, int outtype,
                                      const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                      uint32_t* samplenum, void* const* d_out, const size_t* out_cap)
{
    return mix_multi_dev(m, d_in, in_len, intype, outtype, shift_hz_per_block, nblocks, block_bytes, samplerate, samplenum, d_out, out_cap);
}

}  // extern "C"
