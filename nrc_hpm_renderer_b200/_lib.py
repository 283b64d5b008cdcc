"""ctypes binding of libnrchpm_b200.so -- the C ABI declared in include/nrc_hpm_b200.h.

The library holds sm_100a code only and there is no CPU or PyTorch fallback: if it is missing, or a call
fails, an exception is raised (the reference throws std::runtime_error, src/Log.cpp:16-21).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NRCHPM_LIB selects an experimental build of the same library (scripts/build_variant.sh); development only
LIB_PATH = os.environ.get("NRCHPM_LIB") or os.path.join(_HERE, "libnrchpm_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3


class NrcHpmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[nrchpm error {code}] {msg}")
        self.code = code


class SceneDesc(C.Structure):
    _fields_ = [("dim", C.c_int32 * 3), ("sky_size", C.c_float * 3), ("density_factor", C.c_float), ("g", C.c_float),
                ("dir_light_dir", C.c_float * 3), ("dir_light_strength", C.c_float), ("point_pos", C.c_float * 3),
                ("point_strength", C.c_float), ("point_color", C.c_float * 3), ("env_strength", C.c_float),
                ("env_color", C.c_float * 3)]


class RenderConfig(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("train_width", C.c_uint32), ("train_height", C.c_uint32),
                ("train_x_dist", C.c_uint32), ("train_y_dist", C.c_uint32), ("train_spp", C.c_uint32),
                ("primary_ray_length", C.c_uint32), ("primary_ray_prob", C.c_float), ("train_ring_size", C.c_uint32),
                ("train_ray_length", C.c_uint32), ("infer_batch_size", C.c_uint32), ("blend", C.c_uint32),
                ("show_nrc", C.c_uint32), ("compact_inference", C.c_uint32), ("x_begin", C.c_uint32), ("x_end", C.c_uint32),
                ("train_tx0", C.c_uint32), ("pipeline_train", C.c_uint32)]


class CompareResult(C.Structure):
    _fields_ = [("mse", C.c_float), ("ref_mean", C.c_float), ("own_mean", C.c_float), ("own_var", C.c_float), ("valid_pixel_count", C.c_uint32)]


# name -> (restype, argtypes); every exported symbol of include/nrc_hpm_b200.h is listed here (tests check it)
_P, _F, _U32, _U64, _SZ, _I = C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.c_uint64, C.c_size_t, C.c_int
SIGNATURES = {
    "nrchpm_last_error": (C.c_char_p, []),
    "nrchpm_version": (_I, []),
    "nrchpm_launch_count": (_U64, []),
    "nrc_create": (_I, [C.c_char_p, _U64, C.POINTER(_P)]),
    "nrc_destroy": (_I, [_P]),
    "nrc_init": (_I, [_P, _U32, _P, _P, _P, _P, _P, _P, _P]),
    "nrchpm_import_external_buffer": (_I, [_I, _SZ, C.POINTER(_P), C.POINTER(_P)]),
    "nrchpm_release_external_buffer": (_I, [_P]),
    "nrchpm_import_external_semaphore": (_I, [_I, C.POINTER(_P)]),
    "nrchpm_release_external_semaphore": (_I, [_P]),
    "nrc_infer_and_train": (_I, [_P, C.POINTER(_U32), _I]),
    "nrc_inference": (_I, [_P, C.POINTER(_U32)]),
    "nrc_train": (_I, [_P]),
    "nrc_get_loss": (_I, [_P, _F]),
    "nrc_get_infer_batch_count": (_SZ, [_P]),
    "nrc_get_train_batch_count": (_SZ, [_P]),
    "nrc_get_infer_batch_size": (_U32, [_P]),
    "nrc_get_train_batch_size": (_U32, [_P]),
    "nrc_n_params": (_U64, [_P]),
    "nrc_n_mlp_params": (_U64, [_P]),
    "nrc_input_width": (_U32, [_P]),
    "nrc_get_params": (_I, [_P, _I, _F]),
    "nrc_set_params_fp32": (_I, [_P, _F]),
    "nrc_set_ema": (_I, [_P, _F]),
    "nrc_gradient_buffers": (_I, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "nrc_encode_batch": (_I, [_P, _P, _U32, _I, _P, _P]),
    "nrc_inference_batch": (_I, [_P, _P, _P, _U32, _I, _P]),
    "nrc_inference_indexed": (_I, [_P, _P, _P, _P, _P, _U32, _I, _P]),
    "nrc_snapshot_params": (_I, [_P, _I, _P]),
    "nrc_set_inference_cta_limit": (_I, [_P, _U32]),
    "nrc_peer_export": (_I, [_P, _P]),
    "nrc_peer_setup": (_I, [_P, _I, _I, _P]),
    "nrc_peer_exchange": (_I, [_P, _P]),
    "nrc_training_step": (_I, [_P, _P, _P, _U32, _I, _P]),
    "nrc_optimizer_step": (_I, [_P, _P]),
    "nrc_last_step_tensor": (_I, [_P, _I, _F]),
    "nrc_debug_train_profile": (_U32, [_P, _P, _U32]),
    "nrc_debug_timeline": (_U32, [_P, _P, _U32]),
    "nrc_inference_host": (_I, [_P, _F, _F, _U32, _I]),
    "nrc_training_step_host": (_I, [_P, _F, _F, _U32, _F]),
    "nrc_infer_and_train_host": (_I, [_P, _F, _F, _U32, _F, _F, _U32, _U32, _I, _F]),
    "hpm_scene_create": (_I, [C.POINTER(SceneDesc), _P, C.POINTER(_P)]),
    "hpm_scene_destroy": (_I, [_P]),
    "hpm_renderer_create": (_I, [_P, _P, C.POINTER(RenderConfig), _P, C.POINTER(_P)]),
    "hpm_renderer_destroy": (_I, [_P]),
    "hpm_renderer_set_camera": (_I, [_P, _F, _F]),
    "hpm_renderer_set_blend": (_I, [_P, _I]),
    "hpm_render": (_I, [_P, _F, _I]),
    "hpm_mc_render": (_I, [_P, _F, _U32]),
    "hpm_renderer_set_tracker_mode": (_I, [_P, _I]),
    "hpm_pass_gen_rays": (_I, [_P, _F]),
    "hpm_pass_prep_train": (_I, [_P, _F]),
    "hpm_pass_composite": (_I, [_P]),
    "hpm_sync": (_I, [_P]),
    "hpm_get_stage_ms": (_I, [_P, _F]),
    "hpm_buffer_info": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_SZ)]),
    "hpm_read_buffer": (_I, [_P, _I, _P, _SZ]),
    "hpm_write_buffer": (_I, [_P, _I, _P, _SZ]),
    "hpm_selftest_logf": (_I, [C.POINTER(_U64)]),
    "hpm_compare_images": (_I, [_P, _P, _U32, _U32, C.POINTER(CompareResult), _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library (built in-tree by ``__graft_entry__.build()`` / ``make -C nrc_hpm_renderer_b200/csrc``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NrcHpmError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                        "-- there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise NrcHpmError(rc, lib().nrchpm_last_error().decode("utf-8", "replace"))
