/*
 * doppler_b200.h -- C ABI of libdoppler_b200.so: the B200-native replacement for the per-sample
 * IQ frequency-shift hot path of cubehub/doppler (reference @ 5f13df14, v1.1.10).
 *
 * Drop-in boundary.  The reference's hot path is three Rust functions exported by
 * src/lib.rs:34-35 (`pub mod dsp`) plus the egress casts inlined in src/main.rs:
 *
 *   dsp::convert_iqi16_to_complex(&[u8]) -> Vec<Complex<f32>>              src/dsp.rs:85-99
 *   dsp::convert_iqf32_to_complex(&[u8]) -> Vec<Complex<f32>>              src/dsp.rs:101-115
 *   dsp::shift_frequency(&[Complex<f32>], &mut u32, f32, u32) -> Vec<..>   src/dsp.rs:117-134
 *   (re * 32767.0) as i16 / raw f32 bytes                                  src/main.rs:73-93
 *
 * and its only existing FFI crossing is `extern { fn ccexpf(z: *mut LiquidComplex32); }`
 * (src/dsp.rs:40-42 <- src/complex.c:33): plain C types, in-place, no returned structs.  This
 * header keeps that style: plain pointers and sizes, `int` status returns (0 = ok) where the
 * reference would panic, opaque context.  A Rust `extern "C"` block binds it one-to-one (see
 * INTEGRATION.md).
 *
 * There is NO CPU fallback: every compute entry point runs the sm_100a kernels in
 * doppler_b200/csrc/doppler_b200.cu and returns DOPPLER_B200_ENODEV / ECUDA when it cannot.
 *
 * Numerics contract: output bytes are identical to the reference's for i16 output and
 * bit-identical for f32 output (NaN payloads excepted), on a host whose libm is glibc >= 2.28
 * x86-64 with FMA (the `__sincosf_fma` variant; see doppler_b200/csrc/sincosf_glibc.h).
 */
#ifndef DOPPLER_B200_H
#define DOPPLER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOPPLER_B200_ABI_VERSION 1

/* IQ sample formats: mirrors `enum DataType { F32, I16 }`, src/usage.rs:39-42.
 * I16 = interleaved little-endian int16 I,Q (4 bytes / complex sample);
 * F32 = interleaved little-endian float32 I,Q (8 bytes / complex sample). */
#define DOPPLER_B200_I16 0
#define DOPPLER_B200_F32 1

/* The reference's pump block: `const BUFFER_SIZE: usize = 8192` bytes, src/main.rs:49. */
#define DOPPLER_B200_BUFFER_SIZE 8192

/* Status codes. */
#define DOPPLER_B200_OK 0
#define DOPPLER_B200_EINVAL 1 /* bad argument (null pointer, unknown type, misaligned device pointer) */
#define DOPPLER_B200_EALIGN 2 /* input length not a whole number of samples: the reference's
                                 assert!(len % 4 == 0) / assert!(len % 8 == 0), dsp.rs:87,103 */
#define DOPPLER_B200_ECAP 3   /* output capacity too small */
#define DOPPLER_B200_ECUDA 4  /* CUDA runtime error (see doppler_b200_last_error) */
#define DOPPLER_B200_ENODEV 5 /* no usable sm_100 device / driver */
#define DOPPLER_B200_ENOMEM 6
#define DOPPLER_B200_ELIBM 7  /* DOPPLER_B200_STRICT_LIBM=1 and the host libm's sincosf is not the variant the device
                                 reproduces (doppler_b200_libm_compatible) */

typedef struct doppler_b200_ctx doppler_b200_ctx;

/* ---- lifecycle --------------------------------------------------------------------------- */

int doppler_b200_abi_version(void);

/* One context per GPU (and per caller thread: a context is not re-entrant, like the
 * single-threaded reference).  Owns streams, staging buffers, the phasor-table arena and the
 * samplenum planner cache. */
int doppler_b200_create(int device, doppler_b200_ctx** ctx_out);
void doppler_b200_destroy(doppler_b200_ctx* ctx);

/* Numerics guard.  Output parity with the reference means parity with the HOST libm's sincosf
 * (complex.c:35); the device reproduces glibc >= 2.28 x86-64 `__sincosf_fma`.  Returns 1 when the
 * host twin of the device routine agrees bit for bit with this host's sincosf on ~20 000 probes
 * spanning every range of the algorithm, 0 otherwise (another libm, or glibc's non-FMA variant on a
 * CPU without FMA: the GPU output is then still glibc-FMA-exact, but no longer "what the reference
 * prints on this box").  Evaluated once per process; doppler_b200_create warns on stderr when it is 0
 * and fails with ELIBM under DOPPLER_B200_STRICT_LIBM=1.  Host only, no GPU needed. */
int doppler_b200_libm_compatible(void);
/* Same comparison against any sincosf-shaped function (NULL = the host libm's): number of probes
 * that disagree.  Test hook. */
uint32_t doppler_b200_libm_mismatches(void (*sincosf_fn)(float, float*, float*));

/* Text of the last error on this context ("" if none).  ctx may be NULL for create() errors. */
const char* doppler_b200_last_error(const doppler_b200_ctx* ctx);

/* Pinned host memory (lets the host-buffer entry points DMA directly instead of staging). */
void* doppler_b200_host_alloc(size_t bytes);
void doppler_b200_host_free(void* p);
/* Pins caller-owned, page-aligned host memory in place (cudaHostRegister) so that the host-buffer entry
 * points copy from / to it directly; 0 on success.  The CLI uses it to start reading stdin before the
 * CUDA context exists.  Unregister before freeing the memory. */
int doppler_b200_host_register(void* p, size_t bytes);
int doppler_b200_host_unregister(void* p);

/* Number of kernel launches issued by this context so far (mixer + table builders). */
uint64_t doppler_b200_launch_count(const doppler_b200_ctx* ctx);

/* Thresholds between the library's code paths (results are identical on every path; tests force each path,
 * tools/latency.py tunes them).  SMALL_MAX_SAMPLES: launches up to this many samples take the latency-shaped
 * kernel instead of the persistent bulk-async ones (0 = never).  TINY_HOST_BYTES: host-buffer calls with up to
 * this much input run zero-copy over mapped pinned memory with a completion flag (0 = never; the reference calls
 * its mixer once per 8192-byte block, main.rs:49,70). */
#define DOPPLER_B200_TUNE_SMALL_MAX_SAMPLES 1
#define DOPPLER_B200_TUNE_TINY_HOST_BYTES 2
#define DOPPLER_B200_TUNE_MAX_CLAIM 4   /* most work units a pipeline of the segmented kernels claims at once (1 .. 64) */
#define DOPPLER_B200_TUNE_SEG_VARIANT 3 /* 0: the product's segmented kernels; 1: the experimental 4-warp pipelines where they exist (i16 -> i16) */
#define DOPPLER_B200_TUNE_DECIM_VARIANT 5 /* 0: the fused decimator's register-blocked kernel where the filter fits it; 1: always the generic kernel */
#define DOPPLER_B200_TUNE_DECIM_STAGE_SLOTS 6 /* 8-byte shared-memory slots one CTA step of the register-blocked decimator stages (512 .. 27000) */
#define DOPPLER_B200_TUNE_RESIDENT_IDLE_US 7 /* per-block host calls (up to 32 KiB) are served by a resident one-CTA kernel through a mailbox in pinned host
                                             * memory instead of a launch per block; the kernel leaves after this many microseconds without a request
                                             * (default 20000; 0 = no resident kernel, every block is a launch) */
int doppler_b200_tune(doppler_b200_ctx* ctx, int knob, uint64_t value);

/* ---- the reference's inner boundary, one to one (host buffers) --------------------------- */

/* dsp::convert_iqi16_to_complex, src/dsp.rs:85-99.  out receives len/4 complex f32 samples
 * (2*len/4 floats).  EALIGN where the reference asserts. */
int doppler_b200_convert_iqi16_to_complex(doppler_b200_ctx* ctx, const uint8_t* inbuf, size_t len, float* out);

/* dsp::convert_iqf32_to_complex, src/dsp.rs:101-115 (bit-preserving). */
int doppler_b200_convert_iqf32_to_complex(doppler_b200_ctx* ctx, const uint8_t* inbuf, size_t len, float* out);

/* dsp::shift_frequency, src/dsp.rs:117-134.  inbuf/out: nsamples complex f32 (interleaved).
 * *samplenum is the in/out state the reference keeps at src/main.rs:60. */
int doppler_b200_shift_frequency(doppler_b200_ctx* ctx, const float* inbuf, size_t nsamples, uint32_t* samplenum,
                                 float shift_hz, uint32_t samplerate, float* out);

/* ---- the fused path (replaces convert -> shift_frequency -> egress, main.rs:65-94) -------- */

/* One shift value for the whole buffer (const mode, or one pump block).
 * in: in_len bytes of `intype` samples.  out: receives nsamples * (4|8) bytes of `outtype`.
 * *out_len (optional) = bytes written. */
int doppler_b200_mix(doppler_b200_ctx* ctx, const void* in, size_t in_len, int intype, int outtype, float shift_hz,
                     uint32_t samplerate, uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len);

/* One shift value per `block_bytes` of INPUT (track mode: src/main.rs:177 passes a new
 * shift_hz per BUFFER_SIZE-byte block and carries samplenum, src/main.rs:60).  The last block
 * may be short.  nblocks must be >= ceil(in_len / block_bytes). */
int doppler_b200_mix_blocks(doppler_b200_ctx* ctx, const void* in, size_t in_len, int intype, int outtype,
                            const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                            uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len);

/* ---- fused downstream stage: mix + decimate-by-M FIR (not in the reference) ---------------- */

/* The usual next block after `doppler` in an SDR pipe (README.md:53 feeds a demodulator) is a low-pass
 * decimator; fusing it after the mixer keeps the mixed stream on the chip and shrinks the output -- and for
 * host callers the device -> host copy -- by M.  Definition (oracle/doppler_oracle.c: oracle_mix_decimate):
 *   y[k] = the mixer's Complex<f32> result for stream sample k (the reference's arithmetic, dsp.rs:117-134)
 *   z[m] = sum over t = 0 .. ntaps-1, in that order, of taps[t] * y[m*M - t]   (y = 0 before the stream;
 *          one fused multiply-add in f32 per step, real and imaginary parts separately)
 *   out  = z as f32 pairs, or (z * 32767.0) as i16 like main.rs:73-87
 * Output m is produced by the call that supplies input sample m*M; the object carries the last ntaps-1 mixed
 * samples and the stream position across calls, the caller carries samplenum as everywhere else.  With a
 * per-block schedule, calls must end on block boundaries (as for doppler_b200_mix_blocks). */
typedef struct doppler_b200_decim doppler_b200_decim;
int doppler_b200_decim_create(doppler_b200_ctx* ctx, const float* taps, uint32_t ntaps, uint32_t decimation, doppler_b200_decim** out);
void doppler_b200_decim_destroy(doppler_b200_decim* d);
int doppler_b200_decim_reset(doppler_b200_decim* d);          /* back to the start of a stream (history zeroed) */
uint64_t doppler_b200_decim_position(const doppler_b200_decim* d);   /* input samples consumed so far */
int doppler_b200_mix_decimate(doppler_b200_decim* d, const void* in, size_t in_len, int intype, int outtype, float shift_hz,
                              uint32_t samplerate, uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len);
int doppler_b200_mix_blocks_decimate(doppler_b200_decim* d, const void* in, size_t in_len, int intype, int outtype,
                                     const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                     uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len);
/* device buffers (16-byte aligned), asynchronous on `stream` (NULL = the context's own); one call <= 2^30 samples */
int doppler_b200_mix_decimate_dev(doppler_b200_decim* d, const void* d_in, size_t in_len, int intype, int outtype, float shift_hz,
                                  uint32_t samplerate, uint32_t* samplenum, void* d_out, size_t out_cap, size_t* out_len, void* stream);
int doppler_b200_mix_blocks_decimate_dev(doppler_b200_decim* d, const void* d_in, size_t in_len, int intype, int outtype,
                                         const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                         uint32_t* samplenum, void* d_out, size_t out_cap, size_t* out_len, void* stream);

/* ---- device-resident variants (benchmarks, GPU pipelines, time-sliced multi-GPU) ---------- */

/* d_in / d_out: device pointers, 16-byte aligned, on the context's device.  `stream` is a
 * cudaStream_t passed as void* (NULL = the context's own stream).  Asynchronous: returns after
 * enqueueing; *samplenum is final on return (it is computed analytically on the host). */
int doppler_b200_mix_dev(doppler_b200_ctx* ctx, const void* d_in, size_t in_len, int intype, int outtype, float shift_hz,
                         uint32_t samplerate, uint32_t* samplenum, void* d_out, size_t out_cap, void* stream);

int doppler_b200_mix_blocks_dev(doppler_b200_ctx* ctx, const void* d_in, size_t in_len, int intype, int outtype,
                                const float* shift_hz_per_block, size_t nblocks, size_t block_bytes,
                                uint32_t samplerate, uint32_t* samplenum, void* d_out, size_t out_cap, void* stream);

/* Blocks until everything this context enqueued (on its own stream) has finished. */
int doppler_b200_synchronize(doppler_b200_ctx* ctx);

/* ---- analytic samplenum (host only, no GPU needed) ---------------------------------------- */

/* The reference's samplenum after `count` more samples at a constant shift (dsp.rs:125-130),
 * computed in O(search) instead of O(count): seeds a time slice that starts mid-stream. */
uint32_t doppler_b200_samplenum_advance(uint32_t samplenum, float shift_hz, uint32_t samplerate, uint64_t count);

/* Same for a per-block shift schedule; `count` samples, block_samples samples per block. */
uint32_t doppler_b200_samplenum_advance_blocks(uint32_t samplenum, const float* shift_hz_per_block, size_t nblocks,
                                               uint64_t block_samples, uint32_t samplerate, uint64_t count);

/* ---- time-sliced multi-GPU (one stream over several devices of one box, no collective) ------ */

/* The path shards with no exchange step: the stream is cut into contiguous slices on whole pump
 * blocks (main.rs:49, so the per-block shift of track mode, main.rs:177, stays aligned) and the
 * only cross-slice state -- `samplenum` (main.rs:60) at each slice's first sample -- is carried
 * analytically on the host.  The reference is single-threaded and has no equivalent; these
 * entries make the partition part of the product instead of every caller's job. */

/* [begin, end) in samples of slice `index` of `nslices`: whole blocks of `block_samples`, the
 * remainder blocks to the lowest slices, the ragged tail (a short last block) to the last slice. */
int doppler_b200_slice_bounds(uint64_t total_samples, uint32_t nslices, uint32_t index, uint64_t block_samples,
                              uint64_t* begin, uint64_t* end);

/* begins[i] / seeds[i], i = 0..nslices: first sample of slice i and the reference's samplenum
 * there (begins[nslices] = total_samples, seeds[nslices] = the state after the stream), for a
 * per-block shift schedule; nblocks == 1 means one shift for the whole stream (const mode). */
int doppler_b200_slice_seeds(uint32_t samplenum, const float* shift_hz_per_block, size_t nblocks, uint64_t block_samples,
                             uint32_t samplerate, uint64_t total_samples, uint32_t nslices, uint64_t* begins, uint32_t* seeds);

/* A group of contexts, one per device, each driven by its own host thread.  devices == NULL:
 * ordinals 0..ndevices-1; ndevices == 0: every visible device. */
typedef struct doppler_b200_multi doppler_b200_multi;
int doppler_b200_multi_create(const int* devices, int ndevices, doppler_b200_multi** out);
void doppler_b200_multi_destroy(doppler_b200_multi* m);
int doppler_b200_multi_size(const doppler_b200_multi* m);
doppler_b200_ctx* doppler_b200_multi_ctx(doppler_b200_multi* m, int index);
const char* doppler_b200_multi_last_error(const doppler_b200_multi* m);
uint64_t doppler_b200_multi_launch_count(const doppler_b200_multi* m);

/* doppler_b200_mix / _mix_blocks over all devices of the group: host buffers, slice d through
 * device d's own H2D / kernel / D2H pipeline, all slices concurrently.  Same arguments, same
 * result bytes and the same final *samplenum as the single-device calls. */
int doppler_b200_mix_multi(doppler_b200_multi* m, const void* in, size_t in_len, int intype, int outtype, float shift_hz,
                           uint32_t samplerate, uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len);
int doppler_b200_mix_blocks_multi(doppler_b200_multi* m, const void* in, size_t in_len, int intype, int outtype,
                                  const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                  uint32_t* samplenum, void* out, size_t out_cap, size_t* out_len);

/* Device-resident slices: d_in[d] / d_out[d] live on device d (16-byte aligned), in_len[d] bytes
 * each, consecutive in stream order; with a schedule every slice but the last non-empty one must be
 * whole blocks.  Asynchronous on each context's own stream (doppler_b200_multi_synchronize). */
int doppler_b200_mix_multi_dev(doppler_b200_multi* m, const void* const* d_in, const size_t* in_len, int intype, int outtype,
                               float shift_hz, uint32_t samplerate, uint32_t* samplenum, void* const* d_out, const size_t* out_cap);
int doppler_b200_mix_blocks_multi_dev(doppler_b200_multi* m, const void* const* d_in, const size_t* in_len, int intype, int outtype,
                                      const float* shift_hz_per_block, size_t nblocks, size_t block_bytes, uint32_t samplerate,
                                      uint32_t* samplenum, void* const* d_out, const size_t* out_cap);
int doppler_b200_multi_synchronize(doppler_b200_multi* m);

/* ---- track mode's Doppler schedule (host only, no GPU needed) ------------------------------ */

/* src/main.rs:163: doppler_hz = (range_rate_km_sec * 1000 / c) * frequency * (-1), in f64. */
double doppler_b200_doppler_hz(double range_rate_km_sec, uint32_t frequency);

/* src/main.rs:177: the f32 shift handed to the mixer, `doppler_hz as f32 + offset as f32`. */
float doppler_b200_track_shift(double doppler_hz, int32_t offset);

/* src/main.rs:166: whole seconds of stream time after sample_count samples,
 * `(sample_count as f32 / samplerate as f32) as i64` (f32 arithmetic, saturating cast). */
int64_t doppler_b200_replay_seconds(uint64_t sample_count, uint32_t samplerate);

/* The replay driver's clock, src/main.rs:155-184, with the propagator abstracted as a table:
 * doppler_hz_by_second[s] = the value main.rs:163 yields at start_time + s seconds (index clamped
 * to nsec-1).  Writes the f32 shift the reference uses for every BUFFER_SIZE-byte block of an
 * in_len-byte recording -- block b is evaluated at the whole second reached by the samples
 * counted before block b-1 (one-block lag) -- and returns the number of blocks the reference
 * pumps, in_len / BUFFER_SIZE + 1 (the last one short or empty).  At most `cap` values are
 * written.  Feed the result to doppler_b200_mix_blocks with block_bytes = BUFFER_SIZE. */
size_t doppler_b200_replay_schedule(const double* doppler_hz_by_second, size_t nsec, int32_t offset, uint32_t samplerate,
                                    int intype, size_t in_len, float* shifts_out, size_t cap);

/* ---- orbit propagation for track mode (host only; replaces crate gpredict / libgpredict) --- */

/* Tle::from_file + Predict::new (src/main.rs:141-149): TLE `tlename` from `tlefile`, observer at
 * lat/lon (degrees) and altitude (metres).  SGP4 for near-earth element sets, SDP4 (lunar-solar
 * and resonance terms) for periods >= 225 min, both after Spacetrack Report No. 3 and verified
 * against its test cases; bad checksums are rejected (EINVAL, doppler_b200_tracker_last_error).
 * Parity with libgpredict is UNPINNED (not available offline): track mode from a TLE is functionally
 * equivalent to the reference, not byte-identical -- byte parity is claimed for const mode and for a
 * given Doppler table (doppler_b200_replay_schedule) only. */
typedef struct doppler_b200_tracker doppler_b200_tracker;
int doppler_b200_tracker_create(const char* tlefile, const char* tlename, double lat_deg, double lon_deg, double alt_m,
                                doppler_b200_tracker** out);
int doppler_b200_tracker_create_from_lines(const char* name, const char* line1, const char* line2, double lat_deg, double lon_deg,
                                           double alt_m, doppler_b200_tracker** out);
void doppler_b200_tracker_destroy(doppler_b200_tracker* tr);
/* 1 when the element set takes the deep-space model (SDP4), 0 for SGP4. */
int doppler_b200_tracker_is_deep_space(const doppler_b200_tracker* tr);

/* Physical constants of trackers created afterwards; returns the previous choice (any other argument only
 * queries).  GPREDICT (default; env DOPPLER_B200_ORBIT_CONSTANTS=gpredict|wgs72): the values libgpredict's
 * sgp4sdp4.h carries -- WGS-84 radius / flattening next to the WGS-72 gravity field, rounded qoms2t, s and
 * earth rotation rate -- recalled, not verifiable offline.  WGS72: Spacetrack Report No. 3's own set. */
#define DOPPLER_B200_ORBIT_GPREDICT 0
#define DOPPLER_B200_ORBIT_WGS72 1
int doppler_b200_orbit_constants(int which);
const char* doppler_b200_tracker_last_error(void);

/* Predict::update(Some(t)) (src/main.rs:162) and the fields the reference reads afterwards
 * (main.rs:163,170-173).  unix_seconds is UTC. */
int doppler_b200_tracker_observe(doppler_b200_tracker* tr, double unix_seconds, double* az_deg, double* el_deg, double* range_km,
                                 double* range_rate_km_sec);

/* Bare SGP4 state (TEME, km and km/s) at `minutes_since_epoch`: verification hook. */
int doppler_b200_tracker_teme(doppler_b200_tracker* tr, double minutes_since_epoch, double* pos_km, double* vel_km_s);

/* doppler_hz (main.rs:163) at start, start+1 s, ... : the table doppler_b200_replay_schedule takes. */
size_t doppler_b200_tracker_doppler_table(doppler_b200_tracker* tr, double start_unix_seconds, uint32_t frequency, size_t nsec,
                                          double* doppler_hz_out);

/* Test/introspection hook: expands the planner's closed-form pieces into the per-sample
 * samplenum sequence (trace[k] = value used for sample k) for a per-block schedule.  Returns
 * the number of pieces, *samplenum advanced. */
long doppler_b200_plan_trace(uint32_t* samplenum, const float* shift_hz_per_block, size_t nblocks,
                             uint64_t block_samples, uint32_t samplerate, uint64_t count, uint32_t* trace);

/* Test/introspection hook (host only, no device needed): the work decomposition of ONE kernel
 * launch over `count` samples (count <= 2^30) for a per-block schedule -- the same segment builder
 * and the same tile iterator the sm_100a kernel runs, walked on the host for `npipes` pipelines.
 * For every tile the samplenum of each of its samples is written to trace[k] and cover[k] is
 * incremented (both arrays `count` long, caller-zeroed), so a test can check that every sample
 * below the returned tail start is covered exactly once with the reference's samplenum.
 * stats (optional, 8 words): segments, COLUMN segments, work units, tiles, COLUMN tiles, phasor windows
 * evaluated, samples in COLUMN tiles, samples per full tile.  Returns tail_begin
 * (samples from there to count are mixed one by one from global memory) or -1 on bad arguments. */
long doppler_b200_plan_tiles_trace(int intype, int outtype, uint32_t samplenum, const float* shift_hz_per_block,
                                   size_t nblocks, uint64_t block_samples, uint32_t samplerate, uint64_t count,
                                   uint32_t npipes, uint32_t* trace, uint32_t* cover, uint64_t* stats);

/* Test/introspection hook (host only, no device needed): the register-blocked decimator's walk for a filter of `ntaps` taps
 * and decimation M, for a call whose first output sits at call-relative sample `first_out` -- the same tap layout, segment
 * bounds and run list the sm_100a kernel is launched with, replayed on the host.  One record per walk position u (at most
 * `cap` records, 8 words each): { staged index relative to the thread's base, shared-memory slot relative to the thread's
 * base slot, lowest active output klo, highest active output khi (klo > khi: none), tap index of output 0 .. 3 (0xffffffff
 * where that output is inactive) }; tap_bits[u * 4 + k] (optional, cap * 4 words) = bit pattern of the tap the kernel multiplies
 * for output k at position u.  info (optional, 4 words): threads per CTA that own outputs, staging lead, CTA size, shape.
 * Returns the number of walk positions, 0 when the filter is outside the kernel's envelope, -1 on bad arguments. */
long doppler_b200_decim_walk_trace(const float* taps, uint32_t ntaps, uint32_t decimation, uint64_t first_out, uint32_t* records,
                                   uint32_t* tap_bits, size_t cap, uint32_t* info);

/* Measurement hook: the host-buffer pipeline of doppler_b200_mix (same chunks, slots, streams, staging rules)
 * with the kernel SKIPPED: every chunk goes host -> device (in_len bytes) and the same number of samples'
 * worth of `outtype` bytes comes device -> host (content unspecified).  Its rate is the ceiling the box's
 * host memory / PCIe path sets for doppler_b200_mix on these buffers (bench.py: e2e.ceiling). */
int doppler_b200_pipeline_probe(doppler_b200_ctx* ctx, const void* in, size_t in_len, int intype, int outtype, void* out,
                                size_t out_cap);

/* Device self-test hook: evaluates the kernel's phasor routine, (cos, sin) of
 * theta = (-2*PI) * (r * f32(n)) (dsp.rs:121-122), for n = n0 .. n0+count-1 into host arrays. */
int doppler_b200_phasor_probe(doppler_b200_ctx* ctx, float r, uint32_t n0, size_t count, float* cos_out, float* sin_out);

/* Device self-test hook: the kernel's sincosf on arbitrary float bit patterns
 * first, first+stride, ... (count of them); outputs as host arrays. */
int doppler_b200_sincosf_probe(doppler_b200_ctx* ctx, uint32_t first_bits, uint32_t stride, size_t count,
                               float* sin_out, float* cos_out);

#ifdef __cplusplus
}
#endif
#endif /* DOPPLER_B200_H */
