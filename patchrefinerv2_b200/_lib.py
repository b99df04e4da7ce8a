"""ctypes binding of ``lib/libprv2_b200.so`` (the C ABI declared in include/prv2_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

ABI_VERSION = 206          # == PRV2_ABI_VERSION in include/prv2_b200.h
MAX_SRC = 12
MAX_SEG = 128

EPI_STORE, EPI_LN_GELU, EPI_RESID_F32, EPI_F32, EPI_SHUFFLE, EPI_HEAD = range(6)
ACT_NONE, ACT_RELU, ACT_GELU, ACT_GELU_TANH, ACT_SIGMOID_GATE, ACT_IDENTITY = range(6)


class Prv2Error(RuntimeError):
    pass


class GridStage(C.Structure):
    _fields_ = [("off_h", C.c_int32), ("off_w", C.c_int32), ("n_h", C.c_int32), ("n_w", C.c_int32), ("first", C.c_int32)]


class Src(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("C", C.c_int32), ("cs", C.c_int32)]


class Seg(C.Structure):
    _fields_ = [("src", C.c_int16), ("dh", C.c_int16), ("dw", C.c_int16), ("taps_h", C.c_int16)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cout", C.c_int32),
        ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("block_n", C.c_int32),
        ("n_src", C.c_int32), ("n_seg", C.c_int32),
        ("src", Src * MAX_SRC), ("seg", Seg * MAX_SEG),
        ("weight", C.c_void_p), ("Cout_pad", C.c_int32), ("Ktot", C.c_int32),
        ("epi", C.c_int32), ("act", C.c_int32),
        ("bias", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("eps", C.c_float), ("head_scale", C.c_float),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out_cs", C.c_int32),
        ("relu_hi", C.c_void_p), ("relu_lo", C.c_void_p), ("relu_cs", C.c_int32),
        ("res_hi", C.c_void_p), ("res_lo", C.c_void_p), ("res_cs", C.c_int32),
        ("res2_hi", C.c_void_p), ("res2_lo", C.c_void_p), ("res2_cs", C.c_int32),
        ("out_f32", C.c_void_p), ("out_f32_ld", C.c_int32),
        ("shuffle_k", C.c_int32),
        ("row_map_period", C.c_int32), ("row_map_extra", C.c_int32), ("row_map_offset", C.c_int32),
        ("acc_scale", C.c_float), ("f16", C.c_int32),
    ]


_i, _f, _p, _i64 = C.c_int, C.c_float, C.c_void_p, C.c_int64

#: name -> argtypes (every symbol include/prv2_b200.h declares; tests check the export list)
SIGNATURES = {
    "prv2_version": [],
    "prv2_build_digest": [],
    "prv2_last_error": [],
    "prv2_device_info": [_p],
    "prv2_crop_resize": [_p, _i, _i, _p, _i, _p, _i, _i, _p],
    "prv2_roi_gather_f32": [_p, _i, _i, _i, _p, _i, _f, _p, _p],
    "prv2_roi_gather_act": [_p, _p, _i, _i, _i, _i, _p, _i, _f, _p, _p, _i, _p],
    "prv2_blend_canvas": [_p, _p, _i, _i, _p, _i, _i, _i, _p, _p, _p],
    "prv2_blend_raw": [_p, _p, _i, _i, _p, _p, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p],
    "prv2_blend_raw_prep_bytes": [_i, _i],
    "prv2_blend_raw_prepare": [_p, _i, _i, _i, _p, _p],
    "prv2_blend_partial_canvas": [_p, _p, _p, _i, _i, _p, _i, _i, _i, _p, _p, _p],
    "prv2_blend_partial_raw": [_p, _p, _p, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p, _p],
    "prv2_blend_finalize_canvas": [_p, _p, _p, _i, _i, _p, _i, _i, _i, _p, _p, _p],
    "prv2_blend_finalize_raw": [_p, _p, _i, _i, _p, _p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p],
    "prv2_debug_blend_generic": [_i],
    "prv2_umma_gemm": [_p, _p],
    "prv2_attention": [_p, _p, _i, _i, _i, _p, _p, _p],
    "prv2_layernorm": [_p, _i, _i, _p, _p, _f, _i, _p, _p, _i, _p],
    "prv2_layernorm_gelu": [_p, _i, _i, _p, _p, _f, _p, _p, _i, _p],
    "prv2_patchify": [_p, _i, _i, _i, _p, _p, _i, _p],
    "prv2_assemble_tokens": [_p, _p, _p, _i, _i, _i, _p, _p],
    "prv2_resize_bilinear_act": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _i, _p],
    "prv2_depth_taps": [_p, _p, _i, _i, _i, _p, _p, _i, _i, _i, _p],
    "prv2_tap_stencil": [_p, _i, _i, _i, _i, _p, _p, _p],
    "prv2_final_conv3x3": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p],
    "prv2_nchw_f32_to_act": [_p, _i, _i, _i, _i, _p, _p, _i, _p],
    "prv2_act_to_nchw_f32": [_p, _p, _i, _i, _i, _i, _i, _p, _p],
    "prv2_phase_split": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p],
    "prv2_split_f32": [_p, _i64, _p, _p, _p],
    "prv2_reduce_canvas": [_p, _p, _i64, _p],
    "prv2_dwconv": [_p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _p, _i, _p],
    "prv2_encoder_input": [_p, _p, _i, _i, _i, _p, _p, _p, _p, _i, _p],
    "prv2_zoe_attractor": [_p, _i, _i, _p, _i, _i, _i, _p, _i, _i, _i, _i, _f, _i, _p],
    "prv2_zoe_logbinomial_depth": [_p, _i, _p, _i, _i, _p, _i, _i, _i, _i, _f, _f, _p],
}

_lib = None
#: number of kernel launches issued through the library since the last reset (bench.py's gpu_launches)
launch_count = 0


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load the shared library (no compute).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise Prv2Error(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(the CUDA extension is required; there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_char_p if name in ("prv2_last_error", "prv2_build_digest") else C.c_int64 if name == "prv2_blend_raw_prep_bytes" else C.c_int
    if lib.prv2_version() != ABI_VERSION:
        raise Prv2Error(f"{path} has ABI version {lib.prv2_version()}, this binding needs {ABI_VERSION}: rebuild (python -m patchrefinerv2_b200.build --force)")
    have, want = (lib.prv2_build_digest() or b"").decode(), _build.source_digest()
    if have != want:
        raise Prv2Error(f"{path} was built from other sources (digest {have[:12]} != {want[:12]}): rebuild "
                        "(python -c 'import __graft_entry__ as g; g.build()')")
    _lib = lib
    return lib


#: when set to a list, `call(..., work=(unit, amount))` appends (name, unit, amount, start_event, end_event)
#: so bench.py can time individual kernels with CUDA events on the launching stream
profile_log = None


def call(name: str, *args, work=None) -> None:
    """Call an entry point; raise Prv2Error with the library's message on failure."""
    global launch_count
    lib = load()
    ev = None
    if profile_log is not None and work is not None:
        import torch
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.prv2_last_error()
        raise Prv2Error(f"{name} failed ({rc}): {msg.decode() if msg else ''}")
    if ev is not None:
        ev[1].record()
        profile_log.append((name, work[0], work[1], ev[0], ev[1], work[2] if len(work) > 2 else name))
    launch_count += 1
