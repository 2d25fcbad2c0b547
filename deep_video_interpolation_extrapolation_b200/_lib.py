"""ctypes binding of libflowwarp_b200.so — the C-ABI declared in include/flowwarp_b200.h.

The reference (lzhangbj/deep_video_interpolation_extrapolation) is pure Python over torch and has no
FFI of its own; this stub is the binding a maintainer adds.  There is NO fallback:
if the CUDA library is missing, loading raises and every op of the package fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

FWB_MAX_GROUPS = 4
FWB_PAD_ZEROS, FWB_PAD_BORDER = 0, 1
FWB_FLAG_DETERMINISTIC = 1
FWB_FLAG_ATOMIC_SRC = 2  # grad_src by global atomics (ATen-style), for A/B measurements only
FWB_FLAG_FUSED_BWD = 4  # backward_flow also produces grad_src (fused kernels 2+3, non-deterministic fast path)
FWB_FLAG_GRAD_SRC_ZEROED = 8  # with FUSED_BWD: the caller already zeroed grad_src (e.g. overlapped with the forward)

_f32p = C.POINTER(C.c_float)
i64 = C.c_int64


class fwb_dir(C.Structure):
    _fields_ = [
        ("flow", C.c_void_p), ("flow_sn", i64), ("flow_sc", i64), ("flow_st", i64), ("flow_sh", i64),
        ("gate", C.c_void_p), ("gate_sn", i64), ("gate_st", i64), ("gate_sh", i64),
        ("blend", C.c_void_p), ("blend_sn", i64), ("blend_st", i64), ("blend_sh", i64),
        ("sign", C.c_float), ("_pad", C.c_int32),
    ]


class fwb_group(C.Structure):
    _fields_ = [
        ("C", C.c_int32), ("_pad", C.c_int32),
        ("src", C.c_void_p * 2),
        ("src_sn", i64 * 2), ("src_st", i64 * 2), ("src_sc", i64 * 2), ("src_sh", i64 * 2),
        ("out", C.c_void_p), ("out_sn", i64), ("out_st", i64), ("out_sc", i64), ("out_sh", i64),
    ]


class fwb_problem(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("n_dirs", C.c_int32), ("n_groups", C.c_int32),
        ("padding_mode", C.c_int32), ("align_corners", C.c_int32),
        ("flags", C.c_uint32), ("_pad", C.c_int32),
        ("dir", fwb_dir * 2),
        ("grp", fwb_group * FWB_MAX_GROUPS),
    ]


_G = FWB_MAX_GROUPS


class fwb_grads(C.Structure):
    _fields_ = [
        ("grad_out", C.c_void_p * _G),
        ("go_sn", i64 * _G), ("go_st", i64 * _G), ("go_sc", i64 * _G), ("go_sh", i64 * _G),
        ("grad_src", (C.c_void_p * 2) * _G),
        ("gs_sn", (i64 * 2) * _G), ("gs_st", (i64 * 2) * _G), ("gs_sc", (i64 * 2) * _G), ("gs_sh", (i64 * 2) * _G),
        ("grad_flow", C.c_void_p * 2),
        ("gf_sn", i64 * 2), ("gf_sc", i64 * 2), ("gf_st", i64 * 2), ("gf_sh", i64 * 2),
        ("grad_gate", C.c_void_p * 2),
        ("gg_sn", i64 * 2), ("gg_st", i64 * 2), ("gg_sh", i64 * 2),
        ("grad_blend", C.c_void_p * 2),
        ("gb_sn", i64 * 2), ("gb_st", i64 * 2), ("gb_sh", i64 * 2),
    ]


class fwb_blend(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("T", C.c_int32), ("C", C.c_int32), ("Cn", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("input", C.c_void_p), ("in_sn", i64), ("in_st", i64), ("in_sc", i64), ("in_sh", i64),
        ("mask", C.c_void_p), ("m_sn", i64), ("m_st", i64), ("m_sh", i64),
        ("noise", C.c_void_p), ("nz_sn", i64), ("nz_sc", i64), ("nz_sh", i64),
        ("out", C.c_void_p), ("out_sn", i64), ("out_st", i64), ("out_sc", i64), ("out_sh", i64),
        ("grad_out", C.c_void_p), ("go_sn", i64), ("go_st", i64), ("go_sc", i64), ("go_sh", i64),
        ("grad_input", C.c_void_p), ("gi_sn", i64), ("gi_st", i64), ("gi_sc", i64), ("gi_sh", i64),
        ("grad_mask", C.c_void_p), ("gm_sn", i64), ("gm_st", i64), ("gm_sh", i64),
        ("grad_noise", C.c_void_p), ("gn_sn", i64), ("gn_sc", i64), ("gn_sh", i64),
    ]


class fwb_label_problem(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("T", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("n_dirs", C.c_int32), ("K", C.c_int32), ("padding_mode", C.c_int32), ("align_corners", C.c_int32),
        ("dir", fwb_dir * 2),
        ("labels", C.c_void_p * 2), ("lab_sn", i64 * 2), ("lab_st", i64 * 2), ("lab_sh", i64 * 2),
        ("out", C.c_void_p), ("out_sn", i64), ("out_st", i64), ("out_sc", i64), ("out_sh", i64),
        ("grad_out", C.c_void_p), ("go_sn", i64), ("go_st", i64), ("go_sc", i64), ("go_sh", i64),
        ("grad_flow", C.c_void_p * 2), ("gf_sn", i64 * 2), ("gf_sc", i64 * 2), ("gf_st", i64 * 2), ("gf_sh", i64 * 2),
        ("grad_gate", C.c_void_p * 2), ("gg_sn", i64 * 2), ("gg_st", i64 * 2), ("gg_sh", i64 * 2),
        ("grad_blend", C.c_void_p * 2), ("gb_sn", i64 * 2), ("gb_st", i64 * 2), ("gb_sh", i64 * 2),
        ("accumulate", C.c_int32), ("_pad", C.c_int32),
    ]


class fwb_view(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sn", i64), ("st", i64), ("sc", i64), ("sh", i64)]


LIB_NAME = "libflowwarp_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", LIB_NAME)
if os.environ.get("FWB_LIB"):  # A/B hook: another build of the same library (e.g. different launch bounds)
    LIB_PATH = os.environ["FWB_LIB"]

# every symbol include/flowwarp_b200.h declares
SYMBOLS = (
    "fwb_version",
    "fwb_strerror",
    "fwb_reload_env",
    "fwb_release_cache",
    "fwb_warp_blend_forward",
    "fwb_warp_blend_forward_zero",
    "fwb_label_warp_blend_forward",
    "fwb_label_warp_blend_backward",
    "fwb_mask_blend_forward",
    "fwb_mask_blend_backward",
    "fwb_sample_indices",
    "fwb_workspace_bytes",
    "fwb_warp_blend_backward_flow",
    "fwb_warp_blend_backward_src",
    "fwb_loss_partials_bytes",
    "fwb_flowgrad_loss_forward",
    "fwb_flowgrad_loss_backward",
    "fwb_masked_abs_forward",
    "fwb_masked_abs_backward",
)

_lib = None


class FlowWarpLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library (once).  Raises FlowWarpLibraryError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlowWarpLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C deep_video_interpolation_extrapolation_b200/csrc`.  There is no CPU/torch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    pp, gp, vp = C.POINTER(fwb_problem), C.POINTER(fwb_grads), C.c_void_p
    lib.fwb_version.restype = C.c_int32
    lib.fwb_version.argtypes = []
    lib.fwb_reload_env.restype = None
    lib.fwb_reload_env.argtypes = []
    lib.fwb_release_cache.restype = None
    lib.fwb_release_cache.argtypes = []
    lib.fwb_strerror.restype = C.c_char_p
    lib.fwb_strerror.argtypes = [C.c_int32]
    lib.fwb_warp_blend_forward.restype = C.c_int32
    lib.fwb_warp_blend_forward.argtypes = [pp, vp]
    lib.fwb_warp_blend_forward_zero.restype = C.c_int32
    lib.fwb_warp_blend_forward_zero.argtypes = [pp, gp, vp]
    lp = C.POINTER(fwb_label_problem)
    lib.fwb_label_warp_blend_forward.restype = C.c_int32
    lib.fwb_label_warp_blend_forward.argtypes = [lp, vp]
    lib.fwb_label_warp_blend_backward.restype = C.c_int32
    lib.fwb_label_warp_blend_backward.argtypes = [lp, vp]
    bp = C.POINTER(fwb_blend)
    lib.fwb_mask_blend_forward.restype = C.c_int32
    lib.fwb_mask_blend_forward.argtypes = [bp, vp]
    lib.fwb_mask_blend_backward.restype = C.c_int32
    lib.fwb_mask_blend_backward.argtypes = [bp, vp]
    lib.fwb_sample_indices.restype = C.c_int32
    lib.fwb_sample_indices.argtypes = [pp, C.c_int32, vp, vp, vp, vp, vp, vp]
    lib.fwb_workspace_bytes.restype = C.c_size_t
    lib.fwb_workspace_bytes.argtypes = [pp]
    lib.fwb_warp_blend_backward_flow.restype = C.c_int32
    lib.fwb_warp_blend_backward_flow.argtypes = [pp, gp, vp, C.c_size_t, vp]
    lib.fwb_warp_blend_backward_src.restype = C.c_int32
    lib.fwb_warp_blend_backward_src.argtypes = [pp, gp, vp, C.c_size_t, vp]
    wp = C.POINTER(fwb_view)
    i32 = C.c_int32
    lib.fwb_loss_partials_bytes.restype = C.c_size_t
    lib.fwb_loss_partials_bytes.argtypes = [i32, i32, i32]
    lib.fwb_flowgrad_loss_forward.restype = i32
    lib.fwb_flowgrad_loss_forward.argtypes = [wp, wp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.fwb_flowgrad_loss_backward.restype = i32
    lib.fwb_flowgrad_loss_backward.argtypes = [wp, wp, i32, i32, i32, i32, i32, vp, wp, vp]
    lib.fwb_masked_abs_forward.restype = i32
    lib.fwb_masked_abs_forward.argtypes = [wp, wp, wp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.fwb_masked_abs_backward.restype = i32
    lib.fwb_masked_abs_backward.argtypes = [wp, wp, wp, i32, i32, i32, i32, i32, vp, wp, wp, wp, vp]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    """Map a library return code to the exception class the reference's torch checks raise."""
    if rc == 0:
        return
    msg = load().fwb_strerror(rc).decode()
    if rc < 0:
        raise ValueError(f"{what}: {msg} (code {rc})")
    raise RuntimeError(f"{what}: CUDA error {rc}: {msg}")
