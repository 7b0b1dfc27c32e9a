"""ctypes binding of libposegen_b200.so (include/posegen_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises —
there is no Python/CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POSEGEN_B200_LIB") or os.path.join(_HERE, "lib", "libposegen_b200.so")   # (override: experiment builds)

PGN_OK = 0
PGN_E_INVALID, PGN_E_CUDA, PGN_E_STATE, PGN_E_KERNEL = -1, -2, -3, -4
PRECISION_FP32, PRECISION_BF16 = 0, 1
N_LINEAR = 12

# every symbol include/posegen_b200.h declares (tests check the .so exports all of them)
EXPORTED = [
    "pgn_abi_version", "pgn_last_error", "pgn_create", "pgn_destroy", "pgn_upload_weights",
    "pgn_set_embed_scalars", "pgn_workspace_bytes", "pgn_render_forward", "pgn_activation_dump_bytes",
    "pgn_render_forward_train", "pgn_launch_count",
    "pgn_check_device_status", "pgn_device_status_ptr", "pgn_near_far", "pgn_encode", "pgn_mlp", "pgn_composite", "pgn_composite_backward", "pgn_encode_backward", "pgn_encode_bf16", "pgn_encode_backward_bf16", "pgn_mlp_delta", "pgn_mlp_delta_chain", "pgn_mlp_delta_chain_net", "pgn_mask_dump_bytes", "pgn_render_forward_masks", "pgn_view_delta_from_mask",
    "pgn_sample_pdf", "pgn_generate_rays", "pgn_compose_frame", "pgn_pose_to_skts", "pgn_frame_to_hmr_input",
    "pgn_weight_grad_floats", "pgn_mlp_weight_grads", "pgn_debug_wgrad", "pgn_framecode_backward", "pgn_mlp_input_grads",
    "pgn_pose_fk_backward", "pgn_cylinder_bboxes", "pgn_generate_rays_batch", "pgn_compose_frames_batch", "pgn_gather_ray_rows",
    "pgn_debug_umma_gemm", "pgn_debug_phase_timers",
]


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_joints", "n_samples", "n_importance", "multires", "multires_views",
                                         "net_depth", "net_width", "skip_layer", "device", "n_framecodes", "framecode_ch")]


class NetWeights(C.Structure):
    _fields_ = [("weight", C.c_void_p * N_LINEAR), ("bias", C.c_void_p * N_LINEAR), ("framecodes", C.c_void_p)]


class RenderInputs(C.Structure):
    _fields_ = [("ray_batch", C.c_void_p), ("n_rays", C.c_int64), ("skts", C.c_void_p), ("skts_stride", C.c_int64),
                ("cyls", C.c_void_p), ("cyls_stride", C.c_int64), ("pose_idx", C.c_void_p),
                ("nanfill_chunk", C.c_int64), ("precision", C.c_int32), ("chunk_starts", C.c_void_p), ("n_chunks", C.c_int64),
                ("cams", C.c_void_p), ("lindisp", C.c_int32)]


class RenderOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("rgb_map", "disp_map", "acc_map", "alpha", "rgb0", "disp0", "acc0", "alpha0",
                                          "z_samples", "z_fine", "pdf_inds", "weights0", "raw0", "raw", "near_far")]


class TrainRandom(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("t_rand", "u_is", "noise0", "noise")]


class PosegenError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"posegen_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library; raises if it has not been built (python -m posegen_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -m posegen_b200.build` "
                          "(posegen_b200 has no fallback path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.pgn_abi_version.restype = C.c_int
    lib.pgn_last_error.restype = C.c_char_p
    lib.pgn_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.pgn_destroy.argtypes = [vp]
    lib.pgn_destroy.restype = None
    lib.pgn_upload_weights.argtypes = [vp, C.c_int, C.POINTER(NetWeights), C.c_int, vp]
    lib.pgn_set_embed_scalars.argtypes = [vp, f32, f32, C.POINTER(f32), C.POINTER(f32), f32, f32]
    lib.pgn_workspace_bytes.argtypes = [vp, i64]
    lib.pgn_workspace_bytes.restype = C.c_size_t
    lib.pgn_render_forward.argtypes = [vp, C.POINTER(RenderInputs), C.POINTER(RenderOutputs), vp, C.c_size_t, vp]
    lib.pgn_activation_dump_bytes.argtypes = [i64, i32]
    lib.pgn_activation_dump_bytes.restype = C.c_size_t
    lib.pgn_render_forward_train.argtypes = [vp, C.POINTER(RenderInputs), C.POINTER(RenderOutputs), vp, vp, C.POINTER(TrainRandom), vp, C.c_size_t, vp]
    lib.pgn_launch_count.argtypes = [vp]
    lib.pgn_launch_count.restype = i64
    lib.pgn_check_device_status.argtypes = [vp]
    lib.pgn_device_status_ptr.argtypes = [vp]
    lib.pgn_device_status_ptr.restype = vp
    lib.pgn_near_far.argtypes = [vp, C.POINTER(RenderInputs), vp, vp]
    lib.pgn_encode.argtypes = [vp, C.POINTER(RenderInputs), vp, i32, vp, vp]
    lib.pgn_mlp.argtypes = [vp, C.c_int, vp, i64, vp, i32, vp]
    lib.pgn_encode_backward_bf16.argtypes = [vp, C.POINTER(RenderInputs), vp, i32, vp, vp, i32, vp, vp]
    lib.pgn_encode_bf16.argtypes = [vp, C.POINTER(RenderInputs), vp, i32, vp, vp]
    lib.pgn_mlp_delta.argtypes = [vp, vp, i32, vp, i64, i32, vp, i32, i32, vp, vp, vp, vp]
    lib.pgn_mlp_delta_chain_net.argtypes = [vp, i32, vp, vp, vp, i64, i64, vp, vp, C.c_uint32, vp]
    lib.pgn_mask_dump_bytes.argtypes = [i64]
    lib.pgn_mask_dump_bytes.restype = C.c_size_t
    lib.pgn_render_forward_masks.argtypes = [vp, C.POINTER(RenderInputs), C.POINTER(RenderOutputs), vp, vp, C.c_size_t, vp]
    lib.pgn_view_delta_from_mask.argtypes = [vp, vp, vp, vp, vp, i64, vp]
    lib.pgn_mlp_delta_chain.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp, vp, vp, C.c_uint32, vp]
    lib.pgn_composite.argtypes = [vp, C.POINTER(RenderInputs), vp, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.pgn_composite_backward.argtypes = [vp, C.POINTER(RenderInputs), vp, vp, i32, vp, vp, vp, vp, vp]
    lib.pgn_encode_backward.argtypes = [vp, C.POINTER(RenderInputs), vp, i32, vp, vp, vp]
    lib.pgn_sample_pdf.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp, vp]
    lib.pgn_generate_rays.argtypes = [vp, i32, i32, f32, C.POINTER(f32), i32, i32, i32, i32, vp, vp]
    lib.pgn_compose_frame.argtypes = [vp, i32, i32, i32, i32, i32, i32, vp, vp, f32, vp, vp]
    lib.pgn_pose_to_skts.argtypes = [vp, vp, C.POINTER(f32), i32, f32, f32, f32, vp, vp, vp, vp, vp]
    lib.pgn_frame_to_hmr_input.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, i32, C.POINTER(f32), C.POINTER(f32), i32, vp, vp]
    lib.pgn_mlp_input_grads.argtypes = [vp, i32, vp, vp, i64, vp, vp, i32, vp]
    lib.pgn_framecode_backward.argtypes = [vp, i32, vp, i64, i32, vp, vp, vp, vp]
    lib.pgn_weight_grad_floats.argtypes = [vp]
    lib.pgn_weight_grad_floats.restype = C.c_size_t
    lib.pgn_mlp_weight_grads.argtypes = [vp, i32, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp, vp]
    lib.pgn_debug_wgrad.argtypes = [vp, vp, i32, i32, vp, i32, i32, i64, vp, i32, i32, i32, vp]
    lib.pgn_pose_fk_backward.argtypes = [vp, vp, C.POINTER(f32), i32, vp, vp, vp, vp]
    lib.pgn_cylinder_bboxes.argtypes = [vp, vp, i32, C.POINTER(C.c_double), i32, i32, f32, vp, vp]
    lib.pgn_generate_rays_batch.argtypes = [vp, i32, i32, f32, C.POINTER(f32), vp, vp, i32, i64, vp, vp, vp]
    lib.pgn_compose_frames_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, vp, f32, vp, vp]
    lib.pgn_gather_ray_rows.argtypes = [vp, vp, vp, vp, i64, i64, i32, i64, i64, vp]
    lib.pgn_debug_umma_gemm.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.pgn_debug_phase_timers.argtypes = [vp, i32, C.POINTER(C.c_uint64)]
    for name in EXPORTED:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("pgn_abi_version",):
            pass
    _lib = lib
    return lib


def check(rc: int):
    if rc != PGN_OK:
        raise PosegenError(rc, load().pgn_last_error().decode("utf-8", "replace"))
