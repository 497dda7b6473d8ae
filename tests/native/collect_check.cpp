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
