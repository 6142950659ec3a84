"""ctypes binding of libgmf_b200.so (include/gmf_b200.h).  There is NO CPU fallback: if the CUDA
library is missing or the device is not sm_100, every op raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgmf_b200.so")

_lib: Optional[C.CDLL] = None

EXPORTS = [
    "gmf_last_error", "gmf_version", "gmf_create", "gmf_destroy", "gmf_weight_count", "gmf_weight_spec",
    "gmf_load_weights", "gmf_workspace_bytes", "gmf_pointdsc_forward", "gmf_pointdsc_forward_host",
    "gmf_pointdsc_forward_host_async", "gmf_stream_synchronize",
    "gmf_fusion_layer", "gmf_sc_attention", "gmf_encoder_layer", "gmf_classify", "gmf_pick_seeds",
    "gmf_seed_hypotheses", "gmf_score_hypotheses", "gmf_rigid_transform_3d", "gmf_launch_count",
    "gmf_debug_linear", "gmf_debug_attention", "gmf_profile_enable", "gmf_profile_read",
    "gmf_dgr_head_create", "gmf_dgr_head_destroy", "gmf_dgr_head_weight_count", "gmf_dgr_head_weight_spec",
    "gmf_dgr_head_load_weights", "gmf_dgr_head_forward", "gmf_match_workspace_bytes", "gmf_build_correspondences",
    "gmf_feature_compat_workspace_bytes", "gmf_feature_compat", "gmf_weighted_procrustes", "gmf_debug_plan_host_chunks",
    "gmf_sm_workspace_bytes", "gmf_sm_baseline", "gmf_global_registration",
    "gmf_dgr_head_train_workspace_bytes", "gmf_dgr_head_param_count", "gmf_dgr_head_train_forward", "gmf_dgr_head_train_backward", "gmf_sgd_step",
    "gmf_pointdsc_train_workspace_bytes", "gmf_pointdsc_param_count", "gmf_pointdsc_train_forward", "gmf_pointdsc_train_backward", "gmf_adam_step",
]


class GmfConfig(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("num_iterations", C.c_int32), ("k", C.c_int32), ("ratio", C.c_double),
                ("inlier_threshold", C.c_float), ("nms_radius", C.c_float)]


class GmfError(RuntimeError):
    pass


def load(build_if_missing: bool = False) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        if build_if_missing:
            from .build import build
            build()
        else:
            raise GmfError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(gmf_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.gmf_last_error.restype = C.c_char_p
    lib.gmf_version.restype = C.c_char_p
    lib.gmf_workspace_bytes.restype = C.c_size_t
    lib.gmf_workspace_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.gmf_launch_count.restype = C.c_int64
    lib.gmf_launch_count.argtypes = [C.c_int]
    lib.gmf_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(GmfConfig)]
    lib.gmf_destroy.argtypes = [C.c_void_p]
    lib.gmf_destroy.restype = None
    lib.gmf_weight_count.argtypes = [C.c_int]
    lib.gmf_weight_spec.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
    lib.gmf_load_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    lib.gmf_pointdsc_forward.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.gmf_pointdsc_forward_host.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp]
    lib.gmf_pointdsc_forward_host_async.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp]
    lib.gmf_stream_synchronize.argtypes = [vp, vp]
    lib.gmf_fusion_layer.argtypes = [vp, i, vp, vp, i, i, i, vp, vp, sz, vp]
    lib.gmf_sc_attention.argtypes = [vp, i, vp, vp, vp, i, i, vp, vp, sz, vp]
    lib.gmf_encoder_layer.argtypes = [vp, i, vp, vp, vp, vp, i, i, i, vp, vp, sz, vp]
    lib.gmf_classify.argtypes = [vp, vp, i, i, vp, vp, vp]
    lib.gmf_pick_seeds.argtypes = [vp, vp, vp, i, i, i, vp, vp, sz, vp]
    lib.gmf_seed_hypotheses.argtypes = [vp, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp, sz, vp]
    lib.gmf_score_hypotheses.argtypes = [vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.gmf_rigid_transform_3d.argtypes = [vp, vp, vp, vp, i, i, vp, vp]
    lib.gmf_debug_linear.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, vp, vp]
    lib.gmf_debug_attention.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, f, f, vp, vp, sz, vp]
    lib.gmf_profile_enable.argtypes = [vp, i]
    lib.gmf_profile_read.argtypes = [vp, i, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.gmf_dgr_head_create.argtypes = [C.POINTER(C.c_void_p), i, i, i, i, i]
    lib.gmf_dgr_head_destroy.argtypes = [vp]
    lib.gmf_dgr_head_destroy.restype = None
    lib.gmf_dgr_head_weight_count.argtypes = [i]
    lib.gmf_dgr_head_weight_spec.argtypes = [i, i, C.c_char_p, i, C.POINTER(C.c_int64)]
    lib.gmf_dgr_head_load_weights.argtypes = [vp, vp, C.c_int64]
    lib.gmf_dgr_head_forward.argtypes = [vp, vp, vp, i, i, vp, vp]
    lib.gmf_feature_compat_workspace_bytes.restype = C.c_size_t
    lib.gmf_feature_compat_workspace_bytes.argtypes = [i, i]
    lib.gmf_feature_compat.argtypes = [vp, vp, i, i, vp, vp, sz, vp]
    lib.gmf_debug_plan_host_chunks.argtypes = [i, i, i, i, C.POINTER(C.c_int), i]
    lib.gmf_weighted_procrustes.argtypes = [vp, vp, vp, vp, i, i, f, vp, vp, vp]
    lib.gmf_global_registration.argtypes = [vp, vp, vp, vp, i, i, f, i, i, f, vp, vp, vp, vp]
    lib.gmf_dgr_head_train_workspace_bytes.restype = C.c_size_t
    lib.gmf_dgr_head_train_workspace_bytes.argtypes = [i, i]
    lib.gmf_dgr_head_param_count.restype = C.c_int64
    lib.gmf_dgr_head_param_count.argtypes = [i]
    lib.gmf_dgr_head_train_forward.argtypes = [vp, vp, vp, vp, i, i, vp, vp, sz, vp]
    lib.gmf_dgr_head_train_backward.argtypes = [vp, vp, vp, vp, vp, i, i, vp, vp, vp, vp, sz, vp]
    lib.gmf_sgd_step.argtypes = [vp, vp, vp, C.c_int64, f, f, f, f, i, vp]
    lib.gmf_pointdsc_train_workspace_bytes.restype = C.c_size_t
    lib.gmf_pointdsc_train_workspace_bytes.argtypes = [i, i, i, i, i]
    lib.gmf_pointdsc_param_count.restype = C.c_int64
    lib.gmf_pointdsc_param_count.argtypes = [i]
    lib.gmf_pointdsc_train_forward.argtypes = [i, i, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, f, f, i, vp, vp, vp, vp, vp, sz, vp]
    lib.gmf_pointdsc_train_backward.argtypes = [i, i, vp, vp, vp, vp, i, i, i, f, f, i, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.gmf_adam_step.argtypes = [vp, vp, vp, vp, vp, C.c_int64, f, f, f, f, f, f, i, vp]
    lib.gmf_sm_workspace_bytes.restype = C.c_size_t
    lib.gmf_sm_workspace_bytes.argtypes = [i, i, C.c_double]
    lib.gmf_sm_baseline.argtypes = [vp, vp, vp, i, i, f, C.c_double, i, vp, vp, vp, vp, sz, vp]
    lib.gmf_match_workspace_bytes.restype = C.c_size_t
    lib.gmf_match_workspace_bytes.argtypes = [i, i, i, i]
    lib.gmf_build_correspondences.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise GmfError(f"gmf_b200 error {rc}: {load().gmf_last_error().decode()}")
