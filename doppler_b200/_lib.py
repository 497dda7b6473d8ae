"""ctypes loader for libdoppler_b200.so.  Fails loudly: there is no fallback implementation."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdoppler_b200.so")
# tuning sweeps build variants of the library next to the product one (tools/tune); never a fallback
if os.environ.get("DOPPLER_B200_LIB"):
    LIB_PATH = os.path.abspath(os.environ["DOPPLER_B200_LIB"])

c_ctx = ctypes.c_void_p
u8p = ctypes.POINTER(ctypes.c_uint8)
u32p = ctypes.POINTER(ctypes.c_uint32)
f32p = ctypes.POINTER(ctypes.c_float)
szp = ctypes.POINTER(ctypes.c_size_t)

# name -> (restype, argtypes): exactly the entry points include/doppler_b200.h declares
SIGNATURES = {
    "doppler_b200_abi_version": (ctypes.c_int, []),
    "doppler_b200_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_ctx)]),
    "doppler_b200_destroy": (None, [c_ctx]),
    "doppler_b200_last_error": (ctypes.c_char_p, [c_ctx]),
    "doppler_b200_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "doppler_b200_host_free": (None, [ctypes.c_void_p]),
    "doppler_b200_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    "doppler_b200_host_unregister": (ctypes.c_int, [ctypes.c_void_p]),
    "doppler_b200_libm_compatible": (ctypes.c_int, []),
    "doppler_b200_libm_mismatches": (ctypes.c_uint32, [ctypes.c_void_p]),
    "doppler_b200_tune": (ctypes.c_int, [c_ctx, ctypes.c_int, ctypes.c_uint64]),
    "doppler_b200_launch_count": (ctypes.c_uint64, [c_ctx]),
    "doppler_b200_convert_iqi16_to_complex": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "doppler_b200_convert_iqf32_to_complex": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "doppler_b200_shift_frequency": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_size_t, u32p, ctypes.c_float,
                                                    ctypes.c_uint32, ctypes.c_void_p]),
    "doppler_b200_mix": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_float, ctypes.c_uint32, u32p, ctypes.c_void_p, ctypes.c_size_t, szp]),
    "doppler_b200_mix_blocks": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint32, u32p,
                                               ctypes.c_void_p, ctypes.c_size_t, szp]),
    "doppler_b200_decim_create": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_void_p)]),
    "doppler_b200_decim_destroy": (None, [ctypes.c_void_p]),
    "doppler_b200_decim_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "doppler_b200_decim_position": (ctypes.c_uint64, [ctypes.c_void_p]),
    "doppler_b200_mix_decimate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_float, ctypes.c_uint32, u32p, ctypes.c_void_p, ctypes.c_size_t, szp]),
    "doppler_b200_mix_blocks_decimate": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                                        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint32, u32p,
                                                        ctypes.c_void_p, ctypes.c_size_t, szp]),
    "doppler_b200_mix_decimate_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_float, ctypes.c_uint32, u32p, ctypes.c_void_p, ctypes.c_size_t, szp, ctypes.c_void_p]),
    "doppler_b200_mix_blocks_decimate_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                                            ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint32, u32p,
                                                            ctypes.c_void_p, ctypes.c_size_t, szp, ctypes.c_void_p]),
    "doppler_b200_mix_dev": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_float, ctypes.c_uint32, u32p, ctypes.c_void_p, ctypes.c_size_t,
                                            ctypes.c_void_p]),
    "doppler_b200_mix_blocks_dev": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint32,
                                                   u32p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "doppler_b200_synchronize": (ctypes.c_int, [c_ctx]),
    "doppler_b200_samplenum_advance": (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_float, ctypes.c_uint32, ctypes.c_uint64]),
    "doppler_b200_samplenum_advance_blocks": (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                                                ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64]),
    "doppler_b200_slice_bounds": (ctypes.c_int, [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64,
                                                 ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    "doppler_b200_slice_seeds": (ctypes.c_int, [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32,
                                                ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]),
    "doppler_b200_multi_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "doppler_b200_multi_destroy": (None, [ctypes.c_void_p]),
    "doppler_b200_multi_size": (ctypes.c_int, [ctypes.c_void_p]),
    "doppler_b200_multi_ctx": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int]),
    "doppler_b200_multi_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "doppler_b200_multi_launch_count": (ctypes.c_uint64, [ctypes.c_void_p]),
    "doppler_b200_mix_multi": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_float, ctypes.c_uint32, u32p, ctypes.c_void_p, ctypes.c_size_t, szp]),
    "doppler_b200_mix_blocks_multi": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint32, u32p,
                                                     ctypes.c_void_p, ctypes.c_size_t, szp]),
    "doppler_b200_mix_multi_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_float, ctypes.c_uint32, u32p, ctypes.c_void_p, ctypes.c_void_p]),
    "doppler_b200_mix_blocks_multi_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                         ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint32, u32p,
                                                         ctypes.c_void_p, ctypes.c_void_p]),
    "doppler_b200_multi_synchronize": (ctypes.c_int, [ctypes.c_void_p]),
    "doppler_b200_doppler_hz": (ctypes.c_double, [ctypes.c_double, ctypes.c_uint32]),
    "doppler_b200_track_shift": (ctypes.c_float, [ctypes.c_double, ctypes.c_int32]),
    "doppler_b200_replay_seconds": (ctypes.c_int64, [ctypes.c_uint64, ctypes.c_uint32]),
    "doppler_b200_replay_schedule": (ctypes.c_size_t, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int32, ctypes.c_uint32, ctypes.c_int,
                                                       ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]),
    "doppler_b200_tracker_create": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                                    ctypes.POINTER(ctypes.c_void_p)]),
    "doppler_b200_tracker_create_from_lines": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_double,
                                                               ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_void_p)]),
    "doppler_b200_tracker_destroy": (None, [ctypes.c_void_p]),
    "doppler_b200_tracker_is_deep_space": (ctypes.c_int, [ctypes.c_void_p]),
    "doppler_b200_orbit_constants": (ctypes.c_int, [ctypes.c_int]),
    "doppler_b200_tracker_last_error": (ctypes.c_char_p, []),
    "doppler_b200_tracker_observe": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double] + [ctypes.POINTER(ctypes.c_double)] * 4),
    "doppler_b200_tracker_teme": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p]),
    "doppler_b200_tracker_doppler_table": (ctypes.c_size_t, [ctypes.c_void_p, ctypes.c_double, ctypes.c_uint32, ctypes.c_size_t,
                                                             ctypes.c_void_p]),
    "doppler_b200_plan_trace": (ctypes.c_long, [u32p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32,
                                                ctypes.c_uint64, ctypes.c_void_p]),
    "doppler_b200_plan_tiles_trace": (ctypes.c_long, [ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t,
                                                      ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint32,
                                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "doppler_b200_decim_walk_trace": (ctypes.c_long, [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p,
                                                      ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "doppler_b200_pipeline_probe": (ctypes.c_int, [c_ctx, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                   ctypes.c_size_t]),
    "doppler_b200_phasor_probe": (ctypes.c_int, [c_ctx, ctypes.c_float, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_void_p,
                                                 ctypes.c_void_p]),
    "doppler_b200_sincosf_probe": (ctypes.c_int, [c_ctx, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t,
                                                  ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None


def load():
    """Returns the loaded library with typed entry points; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C doppler_b200/csrc`).  doppler_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
