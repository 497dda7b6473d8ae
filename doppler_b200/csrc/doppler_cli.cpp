// doppler_cli.cpp -- `doppler const|track`: the reference's stdin -> stdout CLI over libdoppler_b200.
//
// Keeps the argv contract of /root/reference/src/usage.rs:117-337 (subcommands, flags, defaults,
// exit codes) and the stream semantics of src/main.rs:57-207 (8192-byte blocks, stop on the first
// short read, one shift per block in track mode, samplenum carried), but pumps the stream in large
// chunks: a reader thread fills pinned buffers, the GPU mixes a whole chunk per call
// (doppler_b200_mix / doppler_b200_mix_blocks reproduce the per-block chain internally), a writer
// thread drains.  stdout is byte-identical to the reference's for the same stdin in const mode and for a given
// per-second Doppler table (--doppler-table); with --tlefile the Doppler comes from this build's own SGP4 / SDP4
// (orbit.cpp) instead of libgpredict, so the output is functionally equivalent, not byte-identical.  Chunks are
// ADAPTIVE: the reader dispatches what has arrived (whole blocks) as soon as stdin goes idle for
// kIdleMicros, so a live 1 Msps pipe sees millisecond latency (the reference: one 8 KiB block) while
// a fast producer fills 32 MiB chunks.
//
// Extensions (not in the reference): --device N; --devices A,B,... (every chunk is cut into contiguous time
// slices, one per listed GPU, with analytically carried samplenum -- doppler_b200_mix_multi); --doppler-table FILE (track replay from a text
// file of doppler_hz values, one per second of recording, instead of --tlefile/--tlename/
// --location/--frequency).
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <poll.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/doppler_b200.h"
#include "orbit.h"

namespace {

const char* kVersion = "1.1.10-b200";
constexpr size_t kBlock = DOPPLER_B200_BUFFER_SIZE;
constexpr size_t kChunkBlocks = 4096;   // 32 MiB of input per GPU call, at most
constexpr long kIdleMicros = 500;       // stdin idle for this long: dispatch what has arrived

// src/main.rs:212-233: "<time>.<ms> [<level> <module> <line>]  <msg>" on stderr
void logline(const char* level, int line, const char* fmt, ...)
{
    char msg[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    struct tm tmv;
    localtime_r(&tv.tv_sec, &tmv);
    char ts[32];
    strftime(ts, sizeof ts, "%Y-%m-%dT%H:%M:%S", &tmv);
    fprintf(stderr, "%s.%3d [%-6s %-30s %3d]  %s\n", ts, (int)(tv.tv_usec / 1000), level, "doppler", line, msg);
}
#define INFO(...) logline("INFO", __LINE__, __VA_ARGS__)
#define ERROR(...) logline("ERROR", __LINE__, __VA_ARGS__)

struct Location {
    double lat, lon, alt;
};

struct Args {
    bool track = false;
    uint32_t samplerate = 0;
    int intype = -1, outtype = -1;
    int32_t shift = 0;
    std::string tlefile, tlename, doppler_table;
    bool have_location = false, have_time = false, have_frequency = false;
    Location location{0, 0, 0};
    int64_t start_unix = 0;
    uint32_t frequency = 0;
    int32_t offset = 0;
    int device = 0;
    std::vector<int> devices;   // --devices: time-slice every chunk across these GPUs
};

[[noreturn]] void usage_error(const char* fmt, ...)
{
    char msg[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(msg, sizeof msg, fmt, ap);
    va_end(ap);
    fprintf(stderr, "error: %s\n\nUSAGE:\n    doppler const --samplerate <SAMPLERATE> --intype <INTYPE> --shift <SHIFT> [--outtype <OUTTYPE>]\n"
                    "    doppler track --samplerate <SAMPLERATE> --intype <INTYPE> --tlefile <TLEFILE> --tlename <TLENAME> "
                    "--location <LOCATION> --frequency <FREQUENCY> [--time <TIME>] [--offset <OFFSET>] [--outtype <OUTTYPE>]\n\n"
                    "For more information try --help\n",
            msg);
    exit(1);   // clap exits 1 on argument errors
}

void print_help(const char* sub)
{
    if (!sub) {
        printf("doppler %s\nCompensates IQ data stream doppler shift based on TLE information, also can be used for doing constant "
               "baseband shifting (B200 build: the mixer runs on the GPU)\n\nUSAGE:\n    doppler [SUBCOMMAND]\n\nFLAGS:\n"
               "    -h, --help       Prints help information\n    -V, --version    Prints version information\n\nSUBCOMMANDS:\n"
               "    const    Constant shift mode\n    track    Doppler tracking mode\n",
               kVersion);
    } else if (!strcmp(sub, "const")) {
        printf("doppler-const\nConstant shift mode\n\nUSAGE:\n    doppler const [OPTIONS] --samplerate <SAMPLERATE> --intype <INTYPE> --shift <SHIFT>\n\n"
               "OPTIONS:\n    -i, --intype <INTYPE>            IQ data input type [values: i16, f32]\n"
               "    -o, --outtype <OUTTYPE>          IQ data output type [values: i16, f32]\n"
               "    -s, --samplerate <SAMPLERATE>    IQ data samplerate\n        --shift <SHIFT>              frequency shift in Hz\n"
               "        --device <N>                 CUDA device (default 0)\n"
               "        --devices <A,B,...>          time-slice the stream across these CUDA devices\n");
    } else {
        printf("doppler-track\nDoppler tracking mode\n\nUSAGE:\n    doppler track [OPTIONS] --samplerate <SAMPLERATE> --intype <INTYPE> --tlefile <TLEFILE> "
               "--tlename <TLENAME> --location <LOCATION> --frequency <FREQUENCY>\n\nOPTIONS:\n"
               "        --frequency <FREQUENCY>      Satellite transmitter frequency in Hz\n"
               "    -i, --intype <INTYPE>            IQ data type [values: i16, f32]\n"
               "        --location <LOCATION>        Observer location (lat=<deg>,lon=<deg>,alt=<m>): eg. lat=58.64560,lon=23.15163,alt=8\n"
               "        --offset <OFFSET>            Constant frequency shift in Hz. Can be used to compensate constant offset\n"
               "    -o, --outtype <OUTTYPE>          IQ data output type [values: i16, f32]\n"
               "    -s, --samplerate <SAMPLERATE>    IQ data samplerate\n"
               "        --time <TIME>                Observation start time in UTC Y-m-dTH:M:S: eg. 2015-05-13T14:28:48. If not specified current time is used\n"
               "        --tlefile <TLEFILE>          TLE file: eg. http://www.celestrak.com/NORAD/elements/cubesat.txt\n"
               "        --tlename <TLENAME>          TLE name in TLE file: eg. ESTCUBE 1\n"
               "        --doppler-table <FILE>       (extension) doppler_hz per second of recording, replaces the TLE options\n"
               "        --device <N>                 CUDA device (default 0)\n"
               "        --devices <A,B,...>          time-slice the stream across these CUDA devices (replay / const)\n");
    }
}

int parse_type(const char* flag, const std::string& v)
{
    if (v == "i16") return DOPPLER_B200_I16;
    if (v == "f32") return DOPPLER_B200_F32;
    usage_error("'%s' isn't a valid value for '%s'\n\t[values: i16 f32]", v.c_str(), flag);
}

template <typename T>
T parse_int(const char* flag, const std::string& v, long long lo, long long hi)
{
    errno = 0;
    char* end = nullptr;
    const long long x = strtoll(v.c_str(), &end, 10);
    if (v.empty() || *end != '\0' || errno == ERANGE || x < lo || x > hi)
        usage_error("Invalid value: The argument '%s' isn't a valid value for %s", v.c_str(), flag);   // value_t_or_exit!
    return (T)x;
}

// usage.rs:85-115
bool parse_location(const std::string& s, Location* out, std::string* err)
{
    if (s.find("lat") == std::string::npos || s.find("lon") == std::string::npos || s.find("alt") == std::string::npos) {
        *err = "--location should be defined as: lat=58.64560,lon=23.15163,alt=8";
        return false;
    }
    bool have[3] = {false, false, false};
    double val[3] = {0, 0, 0};
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t comma = s.find(',', pos);
        if (comma == std::string::npos) comma = s.size();
        const std::string part = s.substr(pos, comma - pos);
        pos = comma + 1;
        const size_t eq = part.find('=');
        if (eq == std::string::npos) continue;
        int which = -1;
        if (part.find("lat") != std::string::npos) which = 0;
        else if (part.find("lon") != std::string::npos) which = 1;
        else if (part.find("alt") != std::string::npos) which = 2;
        if (which < 0) continue;
        size_t eq2 = part.find('=', eq + 1);   // split("=").nth(1)
        const std::string num = part.substr(eq + 1, eq2 == std::string::npos ? std::string::npos : eq2 - eq - 1);
        char* end = nullptr;
        errno = 0;
        const double x = strtod(num.c_str(), &end);
        have[which] = !num.empty() && *end == '\0' && errno == 0;
        val[which] = x;
    }
    if (have[0] && have[1] && have[2]) {
        *out = Location{val[0], val[1], val[2]};
        return true;
    }
    *err = s + " isn't a valid value for --location\n\t[use as: lat=58.64560,lon=23.15163,alt=8]";
    return false;
}

// usage.rs:302-314: %Y-%m-%dT%H:%M:%S, UTC
bool parse_time(const std::string& s, int64_t* unix_out)
{
    struct tm tmv;
    memset(&tmv, 0, sizeof tmv);
    const char* end = strptime(s.c_str(), "%Y-%m-%dT%H:%M:%S", &tmv);
    if (!end || *end != '\0') return false;
    *unix_out = (int64_t)timegm(&tmv);
    return true;
}

Args parse_args(int argc, char** argv)
{
    Args a;
    if (argc < 2) {
        INFO("no arguments provided, try with doppler -h");   // usage.rs:330-333
        exit(1);
    }
    const std::string sub = argv[1];
    if (sub == "-h" || sub == "--help") {
        print_help(nullptr);
        exit(0);
    }
    if (sub == "-V" || sub == "--version") {
        printf("doppler %s\n", kVersion);
        exit(0);
    }
    if (sub != "const" && sub != "track") {
        INFO("no arguments provided, try with doppler -h");
        exit(1);
    }
    a.track = sub == "track";
    bool have_shift = false, have_samplerate = false;
    for (int i = 2; i < argc; i++) {
        std::string k = argv[i], v;
        bool has_v = false;
        if (k == "-h" || k == "--help") {
            print_help(sub.c_str());
            exit(0);
        }
        if (k.rfind("--", 0) == 0) {
            const size_t eq = k.find('=');
            if (eq != std::string::npos) {
                v = k.substr(eq + 1);
                k = k.substr(0, eq);
                has_v = true;
            }
        } else if (k.size() > 2 && k[0] == '-' && (k[1] == 's' || k[1] == 'i' || k[1] == 'o')) {
            v = k.substr(k[2] == '=' ? 3 : 2);
            k = k.substr(0, 2);
            has_v = true;
        }
        auto need = [&]() -> std::string {
            if (has_v) return v;
            if (i + 1 >= argc) usage_error("The argument '%s <value>' requires a value but none was supplied", k.c_str());
            return argv[++i];   // AllowLeadingHyphen: the next token is the value even if it starts with '-'
        };
        if (k == "-s" || k == "--samplerate") {
            a.samplerate = parse_int<uint32_t>("SAMPLERATE", need(), 0, 4294967295LL);
            have_samplerate = true;
        }
        else if (k == "-i" || k == "--intype") a.intype = parse_type("--intype <INTYPE>", need());
        else if (k == "-o" || k == "--outtype") a.outtype = parse_type("--outtype <OUTTYPE>", need());
        else if (k == "--device") a.device = parse_int<int>("DEVICE", need(), 0, 1023);
        else if (k == "--devices") {
            const std::string list = need();
            size_t pos = 0;
            while (pos <= list.size()) {
                size_t comma = list.find(',', pos);
                if (comma == std::string::npos) comma = list.size();
                a.devices.push_back(parse_int<int>("DEVICES", list.substr(pos, comma - pos), 0, 1023));
                pos = comma + 1;
            }
        }
        else if (!a.track && k == "--shift") {
            a.shift = parse_int<int32_t>("SHIFT", need(), -2147483648LL, 2147483647LL);
            have_shift = true;
        } else if (a.track && k == "--tlefile") a.tlefile = need();
        else if (a.track && k == "--tlename") a.tlename = need();
        else if (a.track && k == "--doppler-table") a.doppler_table = need();
        else if (a.track && k == "--location") {
            std::string err;
            if (!parse_location(need(), &a.location, &err)) {   // usage.rs:320-327
                ERROR("%s.", err.c_str());
                exit(1);
            }
            a.have_location = true;
        } else if (a.track && k == "--time") {
            if (!parse_time(need(), &a.start_unix)) {           // usage.rs:302-311
                ERROR("Invalid time.");
                ERROR("--time should be defined in Y-m-dTH:M:S format: eg. 2015-05-13T14:28:48");
                exit(1);
            }
            a.have_time = true;
        } else if (a.track && k == "--frequency") {
            a.frequency = parse_int<uint32_t>("FREQUENCY", need(), 0, 4294967295LL);
            a.have_frequency = true;
        } else if (a.track && k == "--offset") a.offset = parse_int<int32_t>("OFFSET", need(), -2147483648LL, 2147483647LL);
        else usage_error("Found argument '%s' which wasn't expected, or isn't valid in this context", argv[i]);
    }
    if (!have_samplerate) usage_error("The following required arguments were not provided:\n    --samplerate <SAMPLERATE>");
    if (a.intype < 0) usage_error("The following required arguments were not provided:\n    --intype <INTYPE>");
    if (a.outtype < 0) a.outtype = a.intype;   // usage.rs:268-270, 294-296
    if (!a.track && !have_shift) usage_error("The following required arguments were not provided:\n    --shift <SHIFT>");
    if (a.track && a.doppler_table.empty()) {
        if (a.tlefile.empty()) usage_error("The following required arguments were not provided:\n    --tlefile <TLEFILE>");
        if (a.tlename.empty()) usage_error("The following required arguments were not provided:\n    --tlename <TLENAME>");
        if (!a.have_location) usage_error("The following required arguments were not provided:\n    --location <LOCATION>");
        if (!a.have_frequency) usage_error("The following required arguments were not provided:\n    --frequency <FREQUENCY>");
    }
    return a;
}

// ---- chunk pump -------------------------------------------------------------------------------
struct Chunk {
    uint8_t* in = nullptr;
    uint8_t* out = nullptr;
    size_t in_len = 0, out_len = 0;
    bool last = false;
};

class Channel {   // single-producer single-consumer hand-off of chunk indices
public:
    void push(int v)
    {
        std::lock_guard<std::mutex> g(m_);
        q_.push_back(v);
        cv_.notify_one();
    }
    int pop()
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return !q_.empty(); });
        const int v = q_.front();
        q_.erase(q_.begin());
        return v;
    }

private:
    std::mutex m_;
    std::condition_variable cv_;
    std::vector<int> q_;
};

size_t read_full(int fd, uint8_t* buf, size_t want)
{
    size_t got = 0;
    while (got < want) {
        const ssize_t r = read(fd, buf + got, want - got);
        if (r == 0) break;
        if (r < 0) {
            if (errno == EINTR) continue;
            fprintf(stderr, "doppler collect error: %s\n", strerror(errno));   // main.rs:63 expect(...)
            exit(101);
        }
        got += (size_t)r;
    }
    return got;
}

// One adaptive chunk: block until one whole block is there (or EOF), keep taking what arrives while stdin
// stays busy (gaps shorter than kIdleMicros) up to `cap`, then complete the last block so that every chunk
// but the stream's last is a whole number of blocks.  *eof is set when the stream ended.
size_t read_chunk(int fd, uint8_t* buf, size_t cap, bool* eof)
{
    size_t got = 0;
    *eof = false;
    auto take = [&](size_t want) {   // one blocking read of at most `want` bytes
        for (;;) {
            const ssize_t r = read(fd, buf + got, want);
            if (r < 0) {
                if (errno == EINTR) continue;
                fprintf(stderr, "doppler collect error: %s\n", strerror(errno));   // main.rs:63 expect(...)
                exit(101);
            }
            if (r == 0) *eof = true;
            got += (size_t)r;
            return;
        }
    };
    while (!*eof && got < kBlock) take(cap - got);
    while (!*eof && got < cap) {
        struct pollfd pfd = {fd, POLLIN, 0};
        struct timespec ts = {0, kIdleMicros * 1000};
        const int pr = ppoll(&pfd, 1, &ts, nullptr);
        if (pr == 0) break;              // idle: the producer is slower than we are
        if (pr < 0 && errno != EINTR) break;
        if (pr > 0) take(cap - got);
    }
    while (!*eof && got % kBlock != 0) take(kBlock - got % kBlock);
    return got;
}

void write_full(int fd, const uint8_t* buf, size_t len)
{
    size_t done = 0;
    while (done < len) {
        const ssize_t w = write(fd, buf + done, len - done);
        if (w < 0) {
            if (errno == EINTR) continue;
            INFO("doppler stdout.write error: %s", strerror(errno));   // main.rs:86,92 then unwrap() panics
            exit(101);
        }
        done += (size_t)w;
    }
}

[[noreturn]] void die_ctx(doppler_b200_ctx* ctx, const char* what, int rc)
{
    ERROR("%s failed (%d): %s", what, rc, doppler_b200_last_error(ctx));
    exit(1);
}

[[noreturn]] void die_multi(doppler_b200_multi* m, const char* what, int rc)
{
    ERROR("%s failed (%d): %s", what, rc, doppler_b200_multi_last_error(m));
    exit(1);
}

// Shift of pump block `b` in replay mode: the Doppler at the whole second reached by the samples
// counted before block b-1 (main.rs:162-166, one-block lag), plus the offset (main.rs:177).
struct ReplayClock {
    uint32_t samplerate;
    size_t block_samples;
    int64_t second_of_block(uint64_t b) const
    {
        return b == 0 ? 0 : doppler_b200_replay_seconds((b - 1) * (uint64_t)block_samples, samplerate);
    }
};

}  // namespace

int main(int argc, char** argv)
{
    Args args = parse_args(argc, argv);
    INFO("doppler %s (B200 build of cubehub/doppler's mixer path)\n\n", kVersion);

    const size_t ibps = args.intype == DOPPLER_B200_I16 ? 4 : 8, obps = args.outtype == DOPPLER_B200_I16 ? 4 : 8;
    const char* tn[2] = {"i16", "f32"};

    // ---- Doppler source for track mode ----
    std::vector<double> table;                  // --doppler-table
    dorbit::Tracker tracker;                    // --tlefile / --tlename / --location
    bool use_tracker = false;
    if (!args.track) {
        INFO("constant shift mode");
        INFO("\tIQ samplerate   : %u", args.samplerate);
        INFO("\tIQ input type   : %s", tn[args.intype]);
        INFO("\tIQ output type  : %s\n", tn[args.outtype]);
        INFO("\tfrequency shift : %d Hz", args.shift);
    } else {
        INFO("tracking mode");
        INFO("\tIQ samplerate   : %u", args.samplerate);
        INFO("\tIQ input type   : %s", tn[args.intype]);
        INFO("\tIQ output type  : %s\n", tn[args.outtype]);
        if (!args.doppler_table.empty()) {
            FILE* f = fopen(args.doppler_table.c_str(), "r");
            if (!f) {
                INFO("cannot open %s: %s", args.doppler_table.c_str(), strerror(errno));
                return 1;
            }
            double v;
            while (fscanf(f, "%lf", &v) == 1) table.push_back(v);
            fclose(f);
            if (table.empty()) {
                INFO("%s holds no doppler values", args.doppler_table.c_str());
                return 1;
            }
            INFO("\tdoppler table   : %s (%zu s)", args.doppler_table.c_str(), table.size());
        } else {
            INFO("\tTLE file        : %s", args.tlefile.c_str());
            INFO("\tTLE name        : %s", args.tlename.c_str());
            INFO("\tlocation        : Location { lat: %g, lon: %g, alt: %g }", args.location.lat, args.location.lon, args.location.alt);
            std::string err;
            if (!tracker.load(args.tlefile, args.tlename, args.location.lat, args.location.lon, args.location.alt, &err)) {
                INFO("%s", err.c_str());   // main.rs:141-147
                return 1;
            }
            use_tracker = true;
            INFO("\tfrequency       : %u Hz", args.frequency);
            INFO("\tpropagator      : built-in %s, %s constants (stands in for libgpredict: equivalent Doppler, not bit-identical output)",
                 tracker.deep_space() ? "SDP4" : "SGP4", tracker.consts().name);
        }
        if (args.have_time) {
            time_t t = (time_t)args.start_unix;
            struct tm tmv;
            gmtime_r(&t, &tmv);
            char ts[40];
            strftime(ts, sizeof ts, "%Y-%m-%dT%H:%M:%SZ", &tmv);
            INFO("\ttime            : %s", ts);
        }
        INFO("\toffset          : %d Hz\n\n\n", args.offset);
    }

    // ---- chunked modes (const, track replay): start draining stdin NOW, into page-aligned buffers that are
    // pinned in place once the CUDA context exists.  Creating the context takes 1-3 s; a live producer
    // (rtl_fm ...) must not find its pipe full for that long.
    const bool chunked = !(args.track && !args.have_time);
    constexpr int kBufs = 3;
    const size_t in_cap = kChunkBlocks * kBlock, out_cap = in_cap / ibps * obps;
    Chunk chunks[kBufs];
    Channel free_q, filled_q, done_q;
    std::thread reader;
    if (chunked) {
        for (Chunk& c : chunks) {
            c.in = (uint8_t*)aligned_alloc(4096, in_cap);
            c.out = (uint8_t*)aligned_alloc(4096, out_cap);
            if (!c.in || !c.out) {
                ERROR("host allocation failed");
                return 1;
            }
        }
#ifdef F_SETPIPE_SZ
        fcntl(0, F_SETPIPE_SZ, 1 << 20);   // fewer, larger reads / writes when stdin / stdout are pipes (best effort)
        fcntl(1, F_SETPIPE_SZ, 1 << 20);
#endif
        for (int i = 0; i < kBufs; i++) free_q.push(i);
        reader = std::thread([&] {
            for (;;) {
                const int i = free_q.pop();
                Chunk& c = chunks[i];
                c.in_len = read_chunk(0, c.in, in_cap, &c.last);   // the stream ends at the first short block (main.rs:98)
                filled_q.push(i);
                if (c.last) break;
            }
        });
    }

    struct timeval t_start;
    gettimeofday(&t_start, nullptr);
    auto since_start_ms = [&] {
        struct timeval t;
        gettimeofday(&t, nullptr);
        return 1e3 * (double)(t.tv_sec - t_start.tv_sec) + 1e-3 * (double)(t.tv_usec - t_start.tv_usec);
    };
    doppler_b200_ctx* ctx = nullptr;
    doppler_b200_multi* multi = nullptr;
    int rc;
    if (chunked && args.devices.size() > 1) {
        rc = doppler_b200_multi_create(args.devices.data(), (int)args.devices.size(), &multi);
        if (rc == DOPPLER_B200_OK) ctx = doppler_b200_multi_ctx(multi, 0);
    } else {
        rc = doppler_b200_create(args.devices.size() == 1 ? args.devices[0] : args.device, &ctx);
    }
    if (rc != DOPPLER_B200_OK) {
        ERROR("doppler_b200_create failed (%d): %s", rc, doppler_b200_last_error(nullptr));
        fflush(stderr);
        _exit(1);   // the reader thread may be blocked in read(2)
    }
    if (multi) INFO("\ttime slices     : %d GPUs per chunk", doppler_b200_multi_size(multi));
    const double ms_create = since_start_ms();

    uint32_t samplenr = 0;   // main.rs:60

    // ---- realtime track mode (main.rs:186-206): block by block, Doppler at the wall clock ----
    if (args.track && !args.have_time) {
        if (!use_tracker) {
            INFO("realtime tracking needs --tlefile/--tlename/--location/--frequency (a --doppler-table is a replay of a recording: give --time)");
            return 1;
        }
        std::vector<uint8_t> in(kBlock), out(kBlock / ibps * obps);
        time_t last_log = time(nullptr);
        for (;;) {
            struct timeval tv;
            gettimeofday(&tv, nullptr);
            const dorbit::Observation ob = tracker.observe((double)tv.tv_sec + 1e-6 * (double)tv.tv_usec);   // predict.update(None)
            const double doppler_hz = doppler_b200_doppler_hz(ob.range_rate_km_s, args.frequency);
            if (time(nullptr) - last_log >= 1) {   // main.rs:191-199
                last_log = time(nullptr);
                char ts[40];
                struct tm gm;
                gmtime_r(&last_log, &gm);
                strftime(ts, sizeof ts, "%Y-%m-%dT%H:%M:%SZ", &gm);   // time::now_utc().rfc3339()
                INFO("time                : %s", ts);
                INFO("az                  : %.2f\xC2\xB0", ob.az_deg);
                INFO("el                  : %.2f\xC2\xB0", ob.el_deg);
                INFO("range               : %.0f km", ob.range_km);
                INFO("range rate          : %.3f km/sec", ob.range_rate_km_s);
                INFO("doppler@%.3f MHz : %.2f Hz\n", (double)((float)args.frequency / 1000000.0f), doppler_hz);
            }
            const size_t got = read_full(0, in.data(), kBlock);
            size_t out_len = 0;
            rc = doppler_b200_mix(ctx, in.data(), got, args.intype, args.outtype, doppler_b200_track_shift(doppler_hz, args.offset),
                                  args.samplerate, &samplenr, out.data(), out.size(), &out_len);
            if (rc == DOPPLER_B200_EALIGN) {
                fprintf(stderr, "thread 'main' panicked at 'assertion failed: inbuf.len() %% %zu == 0', src/dsp.rs\n", ibps);
                return 101;
            }
            if (rc) die_ctx(ctx, "doppler_b200_mix", rc);
            write_full(1, out.data(), out_len);   // + flush per block (main.rs:97): write(2) is unbuffered
            if (got != kBlock) break;
        }
        doppler_b200_destroy(ctx);
        return 0;
    }

    // ---- const mode and track replay: chunked pump, reader / GPU / writer overlapped ----
    // The chunk buffers are pinned in place (the reader may already be filling them) only once a FULL chunk has arrived,
    // i.e. when the producer is fast enough for the copies to matter: pinning 6 x 32 MiB costs tens of milliseconds, more
    // than the whole job when stdin is a second of a 256 ksps stream (BASELINE configs[0]) or a live radio.
    bool pinned = false;
    auto pin_chunks = [&] {
        const double t0 = since_start_ms();
        for (Chunk& c : chunks)
            if (doppler_b200_host_register(c.in, in_cap) != 0 || doppler_b200_host_register(c.out, out_cap) != 0)
                INFO("could not pin the chunk buffers: %s", doppler_b200_last_error(ctx));
        pinned = true;
        if (getenv("DOPPLER_STATS")) fprintf(stderr, "{\"pinned_buffers_ms\": %.1f}\n", since_start_ms() - t0);
    };
    if (getenv("DOPPLER_STATS")) fprintf(stderr, "{\"startup_ms_context\": %.1f}\n", ms_create);

    std::thread writer([&] {
        for (;;) {
            const int i = done_q.pop();
            if (i < 0) break;
            Chunk& c = chunks[i];
            write_full(1, c.out, c.out_len);
            const bool last = c.last;
            free_q.push(i);
            if (last) break;
        }
    });

    const ReplayClock clock{args.samplerate, kBlock / ibps};
    std::vector<float> shifts;
    uint64_t block0 = 0, bytes_in = 0, nchunks = 0;
    struct timeval pump_t0;
    gettimeofday(&pump_t0, nullptr);
    int64_t last_logged_second = 0;
    int exit_code = 0;
    for (;;) {
        const int i = filled_q.pop();
        Chunk& c = chunks[i];
        if (!pinned && c.in_len == in_cap) pin_chunks();
        // The reference asserts len % bps == 0 in the converter of the short final block
        // (dsp.rs:87,103) AFTER every earlier block has been written: mix the whole samples of
        // the full blocks, then report the panic.
        size_t usable = c.in_len;
        bool panic = false;
        if (c.in_len % ibps != 0) {
            usable = c.in_len / kBlock * kBlock;
            panic = true;
        }
        c.out_len = 0;
        bytes_in += c.in_len;
        nchunks++;
        if (!args.track) {
            if (multi) {
                rc = doppler_b200_mix_multi(multi, c.in, usable, args.intype, args.outtype, (float)args.shift, args.samplerate, &samplenr,
                                            c.out, out_cap, &c.out_len);
                if (rc) die_multi(multi, "doppler_b200_mix_multi", rc);
            } else {
                rc = doppler_b200_mix(ctx, c.in, usable, args.intype, args.outtype, (float)args.shift /* main.rs:110 */, args.samplerate,
                                      &samplenr, c.out, out_cap, &c.out_len);
                if (rc) die_ctx(ctx, "doppler_b200_mix", rc);
            }
        } else {
            const size_t nblocks = (usable + kBlock - 1) / kBlock;
            shifts.resize(nblocks ? nblocks : 1);
            for (size_t b = 0; b < nblocks; b++) {
                const int64_t sec = clock.second_of_block(block0 + b);
                double doppler_hz;
                if (use_tracker) {
                    const dorbit::Observation ob = tracker.observe_cached((double)args.start_unix + (double)sec);
                    doppler_hz = doppler_b200_doppler_hz(ob.range_rate_km_s, args.frequency);
                    // main.rs:166-169: dt is advanced BEFORE the telemetry test, so the 5 s cadence and the printed time
                    // follow the updated dt while az / el / range / doppler are those of the update made at the old dt
                    const int64_t sec_new = clock.second_of_block(block0 + b + 1);
                    if (sec_new - last_logged_second >= 5) {   // main.rs:167-175
                        last_logged_second = sec_new;
                        char ts[40];
                        struct tm gm;
                        const time_t tt = (time_t)(args.start_unix + sec_new);
                        gmtime_r(&tt, &gm);
                        strftime(ts, sizeof ts, "%Y-%m-%dT%H:%M:%SZ", &gm);   // (start_time + dt).to_utc().rfc3339()
                        INFO("time                : %s", ts);
                        INFO("az                  : %.2f\xC2\xB0", ob.az_deg);
                        INFO("el                  : %.2f\xC2\xB0", ob.el_deg);
                        INFO("range               : %.0f km", ob.range_km);
                        INFO("range rate          : %.3f km/sec", ob.range_rate_km_s);
                        INFO("doppler@%.3f MHz : %.2f Hz\n", (double)((float)args.frequency / 1000000.0f), doppler_hz);
                    }
                } else {
                    const size_t idx = sec < 0 ? 0 : ((uint64_t)sec < table.size() ? (size_t)sec : table.size() - 1);
                    doppler_hz = table[idx];
                }
                shifts[b] = doppler_b200_track_shift(doppler_hz, args.offset);
            }
            if (usable && multi) {
                rc = doppler_b200_mix_blocks_multi(multi, c.in, usable, args.intype, args.outtype, shifts.data(), nblocks, kBlock,
                                                   args.samplerate, &samplenr, c.out, out_cap, &c.out_len);
                if (rc) die_multi(multi, "doppler_b200_mix_blocks_multi", rc);
            } else if (usable) {
                rc = doppler_b200_mix_blocks(ctx, c.in, usable, args.intype, args.outtype, shifts.data(), nblocks, kBlock, args.samplerate,
                                             &samplenr, c.out, out_cap, &c.out_len);
                if (rc) die_ctx(ctx, "doppler_b200_mix_blocks", rc);
            }
            block0 += nblocks;
        }
        const bool last = c.last;
        done_q.push(i);
        if (panic) {
            exit_code = 101;   // Rust's panic exit status
            break;
        }
        if (last) break;
    }
    reader.join();
    writer.join();
    if (getenv("DOPPLER_STATS")) {   // not in the reference: pump statistics for tools/cli_bench.sh
        struct timeval t1;
        gettimeofday(&t1, nullptr);
        const double sec = (double)(t1.tv_sec - pump_t0.tv_sec) + 1e-6 * (double)(t1.tv_usec - pump_t0.tv_usec);
        fprintf(stderr, "{\"pump_bytes_in\": %llu, \"pump_seconds\": %.6f, \"chunks\": %llu, \"in_MBps\": %.1f}\n",
                (unsigned long long)bytes_in, sec, (unsigned long long)nchunks, sec > 0 ? 1e-6 * (double)bytes_in / sec : 0.0);
    }
    if (exit_code == 101) fprintf(stderr, "thread 'main' panicked at 'assertion failed: inbuf.len() %% %zu == 0', src/dsp.rs\n", ibps);
    for (Chunk& c : chunks) {
        if (pinned) {
            doppler_b200_host_unregister(c.in);
            doppler_b200_host_unregister(c.out);
        }
        free(c.in);
        free(c.out);
    }
    if (multi)
        doppler_b200_multi_destroy(multi);
    else
        doppler_b200_destroy(ctx);
    return exit_code;
}
