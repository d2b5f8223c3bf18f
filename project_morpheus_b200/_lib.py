"""ctypes binding of ``libsnacb.so`` (the C ABI declared in ``include/snacb.h``).

There is no CPU fallback: if the shared library is missing this module raises, and
``snacb_create`` itself fails without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

ABI_VERSION = 1
OK = 0
WIN_OK, WIN_REJECTED, WIN_CODE4096, WIN_EMPTY, WIN_NONFINITE = 0, 1, 2, 3, 4
NOISE_OFF, NOISE_TENSOR, NOISE_PHILOX = 0, 1, 2
PREC_FP32, PREC_FP16, PREC_FP16X3 = 0, 1, 2
FLAG_NO_RU_FUSION = 1
FLAG_NO_CONVT_NOISE_FUSION = 2
FLAG_PERSISTENT_RU = 4
FLAG_TAIL_FUSION = 8
FLAG_NO_PERSISTENT_CONVT = 16
FLAG_FUSE_RU256 = 32
FLAG_NO_BLOCK_FUSION = 64
FLAG_NO_CONVT_NOISE_COMPOSE = 128
NOISE_PER_FRAME = 3360

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsnacb.so")

_f32p = C.POINTER(C.c_float)


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("precision", C.c_int32),
        ("chunk_items", C.c_int32),
        ("trim", C.c_int32),
        ("flags", C.c_int32),
        ("lanes", C.c_int32),
        ("reserved", C.c_int32 * 9),
    ]


class RuWeights(C.Structure):
    _fields_ = [(n, _f32p) for n in ("alpha1", "dw_w", "dw_b", "alpha2", "pw_w", "pw_b")]


class BlockWeights(C.Structure):
    _fields_ = [("alpha", _f32p), ("convt_w", _f32p), ("convt_b", _f32p), ("noise_w", _f32p), ("ru", RuWeights * 3)]


class EncBlockWeights(C.Structure):
    _fields_ = [("ru", RuWeights * 3), ("alpha", _f32p), ("down_w", _f32p), ("down_b", _f32p)]


class EncoderWeights(C.Structure):
    _fields_ = [("in_w", _f32p), ("in_b", _f32p), ("block", EncBlockWeights * 4), ("out_dw_w", _f32p), ("out_dw_b", _f32p),
                ("inproj_w", _f32p * 3), ("inproj_b", _f32p * 3)]


class Weights(C.Structure):
    _fields_ = [
        ("codebook", _f32p * 3),
        ("outproj_w", _f32p * 3),
        ("outproj_b", _f32p * 3),
        ("head_dw_w", _f32p),
        ("head_dw_b", _f32p),
        ("head_pw_w", _f32p),
        ("head_pw_b", _f32p),
        ("block", BlockWeights * 4),
        ("tail_alpha", _f32p),
        ("tail_w", _f32p),
        ("tail_b", _f32p),
    ]


class KernelStat(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_int64), ("ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


# name -> (restype, argtypes); every symbol include/snacb.h declares
_vp, _i32, _i64, _u64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_size_t
SIGNATURES = {
    "snacb_create": (_i32, [C.POINTER(_vp), C.POINTER(Config)]),
    "snacb_destroy": (None, [_vp]),
    "snacb_last_error": (C.c_char_p, [_vp]),
    "snacb_load_weights": (_i32, [_vp, C.POINTER(Weights)]),
    "snacb_load_encoder_weights": (_i32, [_vp, C.POINTER(EncoderWeights)]),
    "snacb_encode": (_i32, [_vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "snacb_workspace_bytes": (_sz, [_vp]),
    "snacb_launch_count": (_i64, [_vp]),
    "snacb_graph_launch_count": (_i64, [_vp]),
    "snacb_deinterleave": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "snacb_deinterleave_raw": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "snacb_decode_windows": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _i64, _u64, _vp, _vp, _vp, _vp]),
    "snacb_decode_windows_host": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _i32, _vp, _i64, _u64, _vp, _vp, _vp, _vp]),
    "snacb_decode_windows_host_submit": (_i32, [_vp, _vp, _i32, _vp, _i32, _i32, _i32, _u64, _vp, _vp, _vp, _vp, C.POINTER(_i32)]),
    "snacb_decode_windows_host_wait": (_i32, [_vp, _i32]),
    "snacb_decode_codes": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _u64, _vp, _vp, _vp]),
    "snacb_fill_noise": (_i32, [_vp, _u64, _vp, _i32, _i32, _vp, _i64, _vp]),
    "snacb_profile_enable": (_i32, [_vp, _i32]),
    "snacb_profile_read": (_i32, [_vp, C.POINTER(KernelStat), _i32]),
    "snacb_set_tap": (_i32, [_vp, _i32, _vp, _sz]),
    "snacb_plan": (_i32, [_i32, _i32, _i32, _i32, C.POINTER(_i32)]),
    "snacb_get_tap_shape": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    # N2: token ingress (host only)
    "snacb_parse_token": (_i32, [C.c_char_p, _i32, _i32, C.POINTER(_i64)]),
    "snacb_ingest_create": (_i32, [C.POINTER(_vp), _i32]),
    "snacb_ingest_destroy": (None, [_vp]),
    "snacb_ingest_reset": (_i32, [_vp, _i32]),
    "snacb_ingest_push": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "snacb_ingest_finish": (_i32, [_vp, _i32]),
    "snacb_ingest_tick": (_i32, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "snacb_ingest_result": (_i32, [_vp, _i32, _vp, _vp]),
    "snacb_ingest_done": (_i32, [_vp, _i32]),
    "snacb_ingest_stat": (_i64, [_vp, _i32]),
    # N3: PCM egress (host only)
    "snacb_riff_header": (_i32, [_i32, _vp]),
    "snacb_stitch_create": (_i32, [C.POINTER(_vp), _i32, C.c_double]),
    "snacb_stitch_destroy": (None, [_vp]),
    "snacb_stitch_push": (_i64, [_vp, _vp, _i64, _i32, _vp, _i64, C.POINTER(_i32), C.POINTER(_i32)]),
    "snacb_stitch_flush": (_i64, [_vp, _vp, _i64]),
    "snacb_stitch_overlap_samples": (_i64, [_vp]),
    "snacb_stitch_bank_create": (_i32, [C.POINTER(_vp), _i32, _i32, C.c_double]),
    "snacb_stitch_bank_destroy": (None, [_vp]),
    "snacb_stitch_bank_reset": (_i32, [_vp, _i32]),
    "snacb_stitch_bank_push": (_i32, [_vp, _i32, _vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp, _vp]),
    # N3 on the GPU: pinned per-stream PCM rings written by the decode tick
    "snacb_egress_create": (_i32, [C.POINTER(_vp), _i32, _i32, _i32, _i32, C.c_double]),
    "snacb_egress_destroy": (None, [_vp]),
    "snacb_egress_last_error": (C.c_char_p, [_vp]),
    "snacb_egress_overlap_samples": (_i64, [_vp]),
    "snacb_egress_ring_base": (_vp, [_vp, _i32]),
    "snacb_egress_push_device": (_i32, [_vp, _i32, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "snacb_egress_sync": (_i32, [_vp, _vp]),
    "snacb_egress_available": (_i64, [_vp, _i32]),
    "snacb_egress_room": (_i64, [_vp, _i32]),
    "snacb_egress_read": (_i64, [_vp, _i32, _vp, _i64]),
    "snacb_egress_flush": (_i32, [_vp, _i32, _vp]),
    "snacb_egress_reset": (_i32, [_vp, _i32, _vp]),
    "snacb_egress_written": (_i64, [_vp, _i32]),
    "snacb_egress_cursors": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    "snacb_decode_windows_to_ring": (_i32, [_vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m project_morpheus_b200.build` "
            "(or __graft_entry__.build()); there is no CPU fallback for the SNAC decode path"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class SnacbError(RuntimeError):
    pass


def check(lib, handle, rc: int, what: str) -> None:
    if rc != OK:
        msg = lib.snacb_last_error(handle)
        raise SnacbError(f"{what} failed ({rc}): {msg.decode() if msg else 'unknown error'}")
