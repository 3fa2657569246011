"""ctypes declarations of the C ABI in include/a2sb_b200.h (prototypes only, no loading policy)."""
from __future__ import annotations

import ctypes as C

OK = 0
ERR_INVALID = -1
ERR_CUDA = -2
ERR_NOLA = -3
KIND_COMPLEX = 0
KIND_MAGPHASE = 1
MIRROR_PEERS, MIRROR_MULTICAST = 1, 2
OP_COMPLEX_TO_MAGPHASE = 0
OP_MAGPHASE_TO_COMPLEX = 1
OP_PHASE_FIX = 2
OP_POWER_SCALE = 3

c_float_p = C.POINTER(C.c_float)


class FwdArgs(C.Structure):
    _fields_ = [
        ("d_wav", C.c_void_p), ("batch", C.c_int64), ("len", C.c_int64), ("wav_stride", C.c_int64),
        ("sample_first", C.c_int64), ("n_local", C.c_int64), ("t_begin", C.c_int64), ("t_end", C.c_int64),
        ("d_out", C.c_void_p), ("out_pitch", C.c_int64), ("out_kind", C.c_int), ("drop_dc", C.c_int), ("power_on", C.c_int),
        ("power", C.c_float), ("eps", C.c_float), ("stream", C.c_void_p), ("wrap_cols", C.c_int64),
    ]


class InvArgs(C.Structure):
    _fields_ = [
        ("d_spec", C.c_void_p), ("batch", C.c_int64), ("n_frames", C.c_int64), ("spec_T", C.c_int64),
        ("spec_t_first", C.c_int64), ("in_kind", C.c_int), ("has_dc", C.c_int), ("phase_fix", C.c_int),
        ("power_on", C.c_int), ("power", C.c_float), ("eps", C.c_float), ("d_wav", C.c_void_p),
        ("wav_stride", C.c_int64), ("out_first", C.c_int64), ("out_count", C.c_int64), ("stream", C.c_void_p),
    ]


class CorruptArgs(C.Structure):
    _fields_ = [
        ("d_out_corrupt", C.c_void_p), ("d_noise", C.c_void_p), ("noise_pitch", C.c_int64), ("row0", C.c_int64),
        ("row1", C.c_int64), ("col0", C.c_int64), ("col1", C.c_int64), ("level", C.c_float),
    ]


class StepArgs(C.Structure):
    _fields_ = [
        ("d_x_t", C.c_void_p), ("d_x_1", C.c_void_p), ("d_mask", C.c_void_p), ("d_noise_post", C.c_void_p),
        ("d_noise_mask", C.c_void_p), ("d_pred_x0", C.c_void_p), ("d_x_next", C.c_void_p), ("std_fwd_t", C.c_float),
        ("mu_x0", C.c_float), ("mu_xt", C.c_float), ("sd_post", C.c_float), ("std_sb", C.c_float),
        ("mask_pred_x0", C.c_int),
    ]


# name -> (restype, argtypes); every symbol the header declares
PROTOTYPES = {
    "a2sb_last_error": (C.c_char_p, []),
    "a2sb_version": (C.c_int, []),
    "a2sb_is_device_build": (C.c_int, []),
    "a2sb_plan_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "a2sb_set_grid_limit": (C.c_int, [C.c_int, C.c_int]),
    "a2sb_plan_destroy": (C.c_int, [C.c_void_p]),
    "a2sb_num_frames": (C.c_int64, [C.c_int64, C.c_int]),
    "a2sb_istft_length": (C.c_int64, [C.c_int64, C.c_int]),
    "a2sb_stft_forward": (C.c_int, [C.c_void_p, C.POINTER(FwdArgs)]),
    "a2sb_istft_inverse": (C.c_int, [C.c_void_p, C.POINTER(InvArgs)]),
    "a2sb_stft_forward_corrupt": (C.c_int, [C.c_void_p, C.POINTER(FwdArgs), C.POINTER(CorruptArgs)]),
    "a2sb_stft_forward_pcm16": (C.c_int, [C.c_void_p, C.POINTER(FwdArgs)]),
    "a2sb_istft_inverse_pcm16": (C.c_int, [C.c_void_p, C.POINTER(InvArgs)]),
    "a2sb_istft_inverse_mirrored": (C.c_int, [C.c_void_p, C.POINTER(InvArgs), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "a2sb_dft_generic_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p]),
    "a2sb_dft_generic_inverse": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int64, C.c_void_p]),
    "a2sb_pointwise": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_float,
                                 C.c_float, C.c_void_p]),
    "a2sb_griffinlim_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_float,
                                         C.c_void_p]),
    "a2sb_wrap_pad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_float,
                                C.c_void_p]),
    "a2sb_segment_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                      C.c_void_p]),
    "a2sb_segment_blend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                     C.c_void_p]),
    "a2sb_segment_blend_window": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                            C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "a2sb_segment_blend_step": (C.c_int, [C.c_void_p, C.POINTER(StepArgs), C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                          C.c_int, C.c_void_p]),
    "a2sb_rect_mask": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_void_p]),
    "a2sb_mask_with_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    "a2sb_mask_fill": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p]),
    "a2sb_mask_fill_padded": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_void_p]),
    "a2sb_zero_segment_windows": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_void_p]),
    "a2sb_roundtrip_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_float, C.c_float, C.c_int]),
    "a2sb_roundtrip_host_pcm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_float, C.c_float, C.c_int]),
    "a2sb_launch_count": (C.c_int64, []),
    "a2sb_tma_launch_count": (C.c_int64, []),
}


class A2SBError(RuntimeError):
    """Raised for every non-zero status of the C ABI (message = a2sb_last_error())."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def bind(lib: C.CDLL) -> C.CDLL:
    """Attach restype/argtypes for every declared symbol; raises AttributeError if one is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


def check(lib: C.CDLL, rc: int) -> None:
    if rc != OK:
        raise A2SBError(rc, lib.a2sb_last_error().decode("utf-8", "replace"))
