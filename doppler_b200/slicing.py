"""Time-slice partitioning of one IQ stream across ranks.

The path shards with no exchange step: contiguous slices aligned to the reference's 8192-byte
pump block (/root/reference/src/main.rs:49) so that the per-block shift schedule of track mode
(main.rs:177) stays aligned, and the only cross-slice state -- the reference's `samplenum`
(main.rs:60) at the first sample of the slice -- is computed analytically on the host.
Host-only logic: no GPU needed (covered by the world_size-2 gloo test).  Thin caller of the C ABI
(doppler_b200_slice_bounds / _slice_seeds / _samplenum_advance*): the partition lives in the
library, where the CLI (--devices) and doppler_b200_mix_multi use the same rule.
"""
from . import dsp

_BPS = {dsp.I16: 4, dsp.F32: 8}


def block_samples(intype):
    """Samples per BUFFER_SIZE-byte pump block."""
    return dsp.BUFFER_SIZE // _BPS[intype]


def slice_bounds(total_samples, world_size, rank, intype):
    """[begin, end) of `rank`'s slice: whole pump blocks, remainder blocks to the lowest ranks,
    the ragged tail (a short last block) to the last rank."""
    return dsp.slice_bounds(total_samples, world_size, rank, block_samples(intype))   # the C entry: one rule for every caller


def seed_const(shift_hz, samplerate, begin):
    """samplenum at stream sample `begin` for a constant shift (const mode)."""
    return dsp.samplenum_advance(0, shift_hz, samplerate, begin)


def seed_blocks(shifts_hz, intype, samplerate, begin):
    """samplenum at stream sample `begin` for a per-block shift schedule (track mode)."""
    return dsp.samplenum_advance_blocks(0, shifts_hz, block_samples(intype), samplerate, begin)


def plan(shifts_hz, intype, samplerate, total_samples, world_size, samplenum=0):
    """(begins, seeds) of all `world_size` slices in one call (doppler_b200_slice_seeds); a scalar shift = const mode."""
    return dsp.slice_seeds(samplenum, shifts_hz, block_samples(intype), samplerate, total_samples, world_size)
