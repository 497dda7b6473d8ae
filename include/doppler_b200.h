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

typedef struct doppler_b200_ctx doppler_b200_ctx;

/* ---- lifecycle --------------------------------------------------------------------------- */

int doppler_b200_abi_version(void);

/* One context per GPU (and per caller thread: a context is not re-entrant, like the
 * single-threaded reference).  Owns streams, staging buffers, the phasor-table arena and the
 * samplenum planner cache. */
int doppler_b200_create(int device, doppler_b200_ctx** ctx_out);
void doppler_b200_destroy(doppler_b200_ctx* ctx);

/* Text of the last error on this context ("" if none).  ctx may be NULL for create() errors. */
const char* doppler_b200_last_error(const doppler_b200_ctx* ctx);

/* Pinned host memory (lets the host-buffer entry points DMA directly instead of staging). */
void* doppler_b200_host_alloc(size_t bytes);
void doppler_b200_host_free(void* p);

/* Number of kernel launches issued by this context so far (mixer + table builders). */
uint64_t doppler_b200_launch_count(const doppler_b200_ctx* ctx);

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

/* Test/introspection hook: expands the planner's closed-form pieces into the per-sample
 * samplenum sequence (trace[k] = value used for sample k) for a per-block schedule.  Returns
 * the number of pieces, *samplenum advanced. */
long doppler_b200_plan_trace(uint32_t* samplenum, const float* shift_hz_per_block, size_t nblocks,
                             uint64_t block_samples, uint32_t samplerate, uint64_t count, uint32_t* trace);

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
