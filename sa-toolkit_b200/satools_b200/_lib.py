"""ctypes binding of libsatools_hifigan.so (C ABI in include/sa_hifigan.h).

There is no fallback: if the shared library is missing or fails to load, importing the
compute path raises.  Build it with ``python __graft_entry__.py build`` (or
``python -m satools_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_HERE, "..", "csrc", "libsatools_hifigan.so"))

MAX_STAGES = 8
MAX_RB = 4

DTYPE_F32, DTYPE_F16, DTYPE_BF16, DTYPE_F64, DTYPE_PCM16 = 0, 1, 2, 3, 4
PRECISION_FP32, PRECISION_FP16, PRECISION_BF16 = 0, 1, 2
PRECISIONS = {"fp32": PRECISION_FP32, "fp16": PRECISION_FP16, "bf16": PRECISION_BF16}


class Cfg(C.Structure):
    _fields_ = [
        ("input_dim", C.c_int32),
        ("initial_channels", C.c_int32),
        ("n_stages", C.c_int32),
        ("upsample_rates", C.c_int32 * MAX_STAGES),
        ("upsample_kernels", C.c_int32 * MAX_STAGES),
        ("n_resblocks", C.c_int32),
        ("resblock_kernels", C.c_int32 * MAX_RB),
        ("n_dilations", C.c_int32),
        ("resblock_dilations", (C.c_int32 * MAX_RB) * MAX_RB),
        ("device", C.c_int32),
    ]


class YaaptParams(C.Structure):
    """sa_yaapt_params (include/sa_yaapt.h): the `_yaapt` options the front end reads, same names and defaults."""
    _fields_ = [(n, C.c_double) for n in ("sr", "frame_length", "frame_space", "f0_min", "f0_max", "fft_length", "bp_low",
                                          "bp_high", "nlfer_thresh1", "shc_numharms", "shc_window", "shc_pwidth",
                                          "shc_maxpeaks", "shc_thresh1", "shc_thresh2", "f0_double", "f0_half", "merit_extra",
                                          "median_value", "dp5_k1", "spec_pitch_min_std", "tda_frame_length", "nccf_thresh1",
                                          "nccf_thresh2", "nccf_maxcands", "nccf_pwidth", "merit_boost", "nlfer_thresh2", "merit_pivot",
                                          "dp_w1", "dp_w2", "dp_w3", "dp_w4")]


# every symbol include/sa_yaapt.h declares
YAAPT_SYMBOLS = {
    "sa_yaapt_last_error": (C.c_char_p, []),
    "sa_yaapt_default_params": (C.c_int, [C.POINTER(YaaptParams)]),
    "sa_yaapt_padded_length": (C.c_int64, [C.POINTER(YaaptParams), C.c_int64]),
    "sa_yaapt_num_frames": (C.c_int64, [C.POINTER(YaaptParams), C.c_int64]),
    "sa_yaapt_frontend_workspace_bytes": (C.c_size_t, [C.POINTER(YaaptParams), C.c_int32, C.c_int64]),
    "sa_yaapt_shc_length": (C.c_int64, [C.POINTER(YaaptParams)]),
    "sa_yaapt_shc_workspace_bytes": (C.c_size_t, [C.POINTER(YaaptParams), C.c_int32, C.c_int64]),
    "sa_yaapt_shc": (C.c_int, [C.POINTER(YaaptParams), C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_yaapt_spec_track_workspace_bytes": (C.c_size_t, [C.POINTER(YaaptParams), C.c_int32, C.c_int64]),
    "sa_yaapt_spec_track": (C.c_int, [C.POINTER(YaaptParams), C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int32),
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_yaapt_track_workspace_bytes": (C.c_size_t, [C.POINTER(YaaptParams), C.c_int32, C.c_int64]),
    "sa_yaapt_track": (C.c_int, [C.POINTER(YaaptParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                 C.c_int64, C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_yaapt_frontend": (C.c_int, [C.POINTER(YaaptParams), C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_int32), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

# name -> (restype, argtypes): every symbol include/sa_hifigan.h declares.
SYMBOLS = {
    "sa_hifigan_abi_version": (C.c_int, []),
    "sa_hifigan_last_error": (C.c_char_p, []),
    "sa_hifigan_default_cfg": (C.c_int, [C.POINTER(Cfg)]),
    "sa_hifigan_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(C.c_void_p)]),
    "sa_hifigan_destroy": (None, [C.c_void_p]),
    "sa_hifigan_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32, C.c_int32]),
    "sa_hifigan_finalize": (C.c_int, [C.c_void_p, C.c_int32]),
    "sa_hifigan_output_length": (C.c_int64, [C.c_void_p, C.c_int64]),
    "sa_hifigan_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32]),
    "sa_hifigan_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_void_p,
                                     C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_hifigan_forward_parts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                           C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t,
                                           C.c_void_p]),
    "sa_hifigan_host_scratch_bytes": (C.c_size_t, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "sa_hifigan_synthesize_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                             C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_hifigan_synthesize_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                                   C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_hifigan_synthesize_host_parts_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                         C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_int32,
                                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_hifigan_set_codebook": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "sa_hifigan_forward_vq": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_int32), C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_hifigan_vq_assign": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sa_hifigan_synthesize_host_trimmed_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32),
                                                           C.c_void_p, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_hifigan_synthesize_host_vq_trimmed_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                                              C.c_int32, C.POINTER(C.c_int32), C.c_void_p, C.c_int32,
                                                              C.c_void_p, C.c_size_t, C.c_void_p]),
    "sa_hifigan_set_debug_tap": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "sa_hifigan_last_launch_count": (C.c_int64, [C.c_void_p]),
    "sa_hifigan_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sa_hifigan_chain_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "sa_hifigan_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "sa_hifigan_get_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int32]),
}

_lib = None


class SaHifiganError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the library once; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SaHifiganError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python __graft_entry__.py build`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in list(SYMBOLS.items()) + list(YAAPT_SYMBOLS.items()):
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.sa_hifigan_abi_version() != 1:
        raise SaHifiganError("libsatools_hifigan.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sa_hifigan_last_error()
        raise SaHifiganError(f"sa_hifigan error {rc}: {msg.decode() if msg else ''}")
