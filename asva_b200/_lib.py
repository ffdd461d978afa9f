"""ctypes binding of libasva_b200.so (the C ABI declared in include/asva_b200.h).

The product path has no CPU fallback: `load()` raises if the shared object is missing or does not export the
full ABI, and every call raises `AsvaError` on a non-zero status."""
import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libasva_b200.so")

MAX_SEG = 10


class AsvaError(RuntimeError):
    pass


class GemmSeg(C.Structure):
    _fields_ = [("src", C.c_int32), ("c0", C.c_int32), ("off", C.c_int32 * 3), ("num_kb", C.c_int32),
                ("wk", C.c_int32), ("wk_first", C.c_int32), ("fix2", C.c_int32), ("reserved", C.c_int32)]


class RowAdd(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int64), ("div", C.c_int32), ("reserved", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p * 2),
        ("a_dims", (C.c_int64 * 4) * 2),
        ("a_strides", (C.c_int64 * 3) * 2),
        ("box", C.c_int32 * 3),
        ("trav", C.c_int32 * 3),
        ("out_dims", C.c_int32 * 3),
        ("nseg", C.c_int32),
        ("seg", GemmSeg * MAX_SEG),
        ("w", C.c_void_p),
        ("ldw", C.c_int64),
        ("N", C.c_int32),
        ("K", C.c_int32),
        ("wcols", C.c_int32),
        ("cta_group", C.c_int32),
        ("bias", C.c_void_p),
        ("add", RowAdd),
        ("res", C.c_void_p * 2),
        ("res_ld", C.c_int64 * 2),
        ("geglu", C.c_int32),
        ("out_fp32", C.c_int32),
        ("out", C.c_void_p),
        ("ldo", C.c_int64),
        ("block_n", C.c_int32),
        ("split_k", C.c_int32),
        ("ws", C.c_void_p),
        ("ws_bytes", C.c_int64),
        ("epilogue", C.c_int32),
        ("ln_cols", C.c_int32),
        ("stats_out", C.c_void_p),
        ("ln_stats", C.c_void_p),
        ("ln_wsum", C.c_void_p),
        ("ln_stat_rows", C.c_int64),
        ("ln_grp_rows", C.c_int32),
        ("ln_grp_stride", C.c_int32),
        ("ln_eps", C.c_float),
        ("reserved", C.c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("kv", C.c_void_p), ("mask", C.c_void_p), ("out", C.c_void_p),
        ("ldq", C.c_int64), ("ldkv", C.c_int64), ("ldo", C.c_int64), ("mask_ld", C.c_int64),
        ("G", C.c_int32), ("heads", C.c_int32), ("R", C.c_int32), ("Nk", C.c_int32), ("d", C.c_int32),
        ("dpad", C.c_int32),
        ("kv_rows_per_group", C.c_int32), ("k_col0", C.c_int32), ("v_col0", C.c_int32), ("mask_rows", C.c_int32),
        ("scale", C.c_float), ("form", C.c_int32),
    ]


# name -> (restype, argtypes); must list every symbol include/asva_b200.h declares
ABI = {
    "asva_gemm": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p]),
    "asva_gemm_plan": (C.c_int, [C.POINTER(GemmDesc), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "asva_gemm_tune": (C.c_int, [C.POINTER(GemmDesc), C.c_void_p, C.c_int32, C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                 C.POINTER(C.c_float)]),
    "asva_attention": (C.c_int, [C.POINTER(AttnDesc), C.c_void_p]),
    "asva_temporal_attention": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_float, C.c_void_p]),
    "asva_temporal_attention_form": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "asva_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                 C.c_float, C.c_int32, C.c_int32, C.c_void_p]),
    "asva_groupnorm_stats": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int64,
                                       C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "asva_groupnorm_ws_floats": (C.c_int64, [C.c_int32, C.c_int64, C.c_int32]),
    "asva_groupnorm": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                C.c_float, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "asva_groupnorm_sync_bytes": (C.c_int64, []),
    "asva_groupnorm_form": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_int32]),
    "asva_groupnorm_apply": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p]),
    "asva_conv_in_im2col": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_void_p]),
    "asva_tconv_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "asva_conv_out_finish": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "asva_small_linear": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "asva_timestep_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "asva_softmax_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_float,
                                    C.c_void_p]),
    "asva_cfg_ddim_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_void_p]),
    "asva_cfg_plms_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "asva_last_error": (C.c_char_p, []),
    "asva_version": (C.c_int, []),
    "asva_device_check": (C.c_int, []),
}

_lib: Optional[C.CDLL] = None


def load(path: Optional[str] = None) -> C.CDLL:
    """Loads the shared object and binds every ABI symbol. Raises AsvaError if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    # ASVA_LIB: an experiment build (tools/build_debug.sh: -DASVA_DEBUG_SWITCHES, traces) instead of the shipped library
    p = path or os.environ.get("ASVA_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise AsvaError(
            f"{p} not found: build it with `python -m asva_b200.build` (or __graft_entry__.build()); "
            "there is no CPU fallback for the CUDA hot path")
    lib = C.CDLL(p)
    for name, (res, args) in ABI.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise AsvaError(f"{p} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().asva_last_error()
        raise AsvaError(f"{what} failed ({status}): {msg.decode() if msg else '?'}")
