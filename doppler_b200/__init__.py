"""doppler_b200 -- B200-native NCO mixer behind cubehub/doppler's own function boundary.

The product is the C-ABI shared library ``libdoppler_b200.so`` (``include/doppler_b200.h``,
sources in ``doppler_b200/csrc``).  This package is the thin Python binding used by the tests
and ``bench.py``; it mirrors the reference's ``dsp`` module (``/root/reference/src/dsp.rs``).
There is no CPU implementation here: every compute call runs the sm_100a kernels or raises.
"""
from . import dsp  # noqa: F401
from .dsp import (  # noqa: F401
    F32,
    I16,
    BUFFER_SIZE,
    DopplerError,
    Mixer,
    MultiMixer,
    Decimator,
    convert_iqf32_to_complex,
    convert_iqi16_to_complex,
    samplenum_advance,
    samplenum_advance_blocks,
    shift_frequency,
)

__all__ = [
    "dsp", "Mixer", "MultiMixer", "Decimator", "DopplerError", "I16", "F32", "BUFFER_SIZE",
    "convert_iqi16_to_complex", "convert_iqf32_to_complex", "shift_frequency",
    "samplenum_advance", "samplenum_advance_blocks",
]
