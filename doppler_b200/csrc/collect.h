// doppler_b200/csrc/collect.h -- host side of the resident kernel's fence-free hand-over (mixer_kernels.cuh: mix_resident_kernel).
// The device writes its result as 8-byte units {word, request number}; a word has arrived when its neighbour shows the number.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace dcollect {
// Request side.  A request is `sectors` 32-byte sectors of 7 payload words + a tag.  `payload` holds at most sectors * 7 words
// (the rest is zero).  Writes every sector's payload before its tag, sector by sector, then drains the store buffer.  The device
// takes the request when every tag shows `seq` (mixer_kernels.cuh: rt_verdict).
void post(volatile uint32_t* mailbox, int sectors, uint32_t seq, const uint32_t* payload, int payload_words);

// Result words [from, n) of request `seq` out of their units into `out` (any alignment), as far as they have arrived.
// Returns the index of the first word that has not (n when the result is complete).
size_t collect(const uint32_t* units, uint32_t seq, unsigned char* out, size_t from, size_t n);
// (tests) the scalar path alone
size_t collect_scalar(const uint32_t* units, uint32_t seq, unsigned char* out, size_t from, size_t n);
}   // namespace dcollect
