"""ctypes mirror of include/vkgs_b200.h and loader of the in-tree CUDA library.

The library is REQUIRED: importing succeeds without it (so CPU-only tooling can import the
package), but every call goes through `lib()` which raises if libvkgs_b200.so is missing —
there is no Python/CPU fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os as _os
# VKGS_LIB: tuning builds only (tools/ab_variants.py); the product library is lib/libvkgs_b200.so
LIB_PATH = Path(_os.environ["VKGS_LIB"]) if _os.environ.get("VKGS_LIB") else PKG / "lib" / "libvkgs_b200.so"

VKGS_OK = 0
VKGS_ERR_INVALID_ARGUMENT = -1
VKGS_ERR_CUDA = -2
VKGS_ERR_NO_DEVICE = -3
VKGS_ERR_NOT_UPLOADED = -4
VKGS_ERR_OVERFLOW = -5
VKGS_ERR_UNSUPPORTED = -6
VKGS_ERR_IO = -7
VKGS_ERR_OUT_OF_MEMORY = -8

FORMAT_FLOAT32, FORMAT_FLOAT16, FORMAT_UINT8 = 0, 1, 2
FRUSTUM_CULLING_NONE, FRUSTUM_CULLING_AT_DIST, FRUSTUM_CULLING_AT_RASTER = 0, 1, 2
SIZE_CULLING_DISABLED, SIZE_CULLING_ENABLED = 0, 1
PIPELINE_3DGS, PIPELINE_3DGUT = 0, 1
EXTENT_EIGEN, EXTENT_CONIC = 0, 1
CAMERA_PINHOLE, CAMERA_FISHEYE = 0, 1

K_NAMES = ["preprocess", "sort_hist", "sort_pass0", "sort_pass1", "sort_pass2", "sort_pass3", "bin_emit",
           "tile_hist", "tile_sort0", "tile_sort1", "tile_ranges", "blend"]
K_COUNT = 12

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)


class SplatSetView(C.Structure):
    _fields_ = [("positions", f32p), ("f_dc", f32p), ("f_rest", f32p), ("opacity", f32p), ("scale", f32p),
                ("rotation", f32p), ("count", C.c_uint64), ("f_rest_per_splat", C.c_uint32), ("_pad", C.c_uint32)]


class Options(C.Structure):
    _fields_ = [("frustum_culling_mode", C.c_uint32), ("size_culling_mode", C.c_uint32), ("front_to_back", C.c_uint32),
                ("ms_antialiasing", C.c_uint32), ("sh_format", C.c_uint32), ("rgba_format", C.c_uint32),
                ("point_cloud_mode", C.c_uint32), ("show_sh_only", C.c_uint32), ("disable_opacity_gaussian", C.c_uint32),
                ("transmittance_epsilon", C.c_float), ("target_format", C.c_uint32), ("surface_info", C.c_uint32),
                ("pipeline", C.c_uint32), ("extent_projection", C.c_uint32), ("kernel_degree", C.c_uint32),
                ("camera_model", C.c_uint32), ("quantize_normals", C.c_uint32), ("_reserved", C.c_uint32 * 1)]


class FrameParams(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("model", C.c_float * 16),
                ("model_inverse", C.c_float * 16), ("camera_position", C.c_float * 3), ("focal", C.c_float * 2),
                ("viewport", C.c_float * 2), ("basis_viewport", C.c_float * 2), ("inverse_focal_adjustment", C.c_float),
                ("splat_scale", C.c_float), ("frustum_dilation", C.c_float), ("alpha_cull_threshold", C.c_float),
                ("size_culling_min_pixels", C.c_float), ("sh_degree", C.c_uint32), ("width", C.c_uint32),
                ("height", C.c_uint32), ("depth_iso_threshold", C.c_float), ("thin_particle_threshold", C.c_float),
                ("view_inverse", C.c_float * 16), ("proj_inverse", C.c_float * 16), ("view_quat", C.c_float * 4),
                ("view_trans", C.c_float * 3), ("near_far", C.c_float * 2), ("alpha_clamp", C.c_float),
                ("kernel_min_response", C.c_float), ("fov_rad", C.c_float)]


class Camera(C.Structure):
    _fields_ = [("eye", C.c_float * 3), ("ctr", C.c_float * 3), ("up", C.c_float * 3), ("fov_deg", C.c_float),
                ("znear", C.c_float), ("zfar", C.c_float)]


class Outputs(C.Structure):
    _fields_ = [("rgba", f32p), ("sorted_ids", u32p), ("sorted_keys", u32p), ("sorted_ids_capacity", C.c_uint64),
                ("visible_count", C.c_uint32), ("_pad", C.c_uint32), ("tile_pairs", C.c_uint64), ("ms_dist", C.c_float),
                ("ms_sort", C.c_float), ("ms_raster", C.c_float), ("ms_total", C.c_float), ("ms_kernel", C.c_float * 16),
                ("bytes_algorithmic", C.c_uint64), ("list_entries_evaluated", C.c_uint64), ("fragments_blended", C.c_uint64)]


class Instance(C.Structure):
    _fields_ = [("splat_set_index", C.c_uint32), ("_pad", C.c_uint32), ("transform", C.c_float * 16),
                ("transform_inverse", C.c_float * 16)]


class ImageMetrics(C.Structure):
    _fields_ = [("mse", C.c_float), ("psnr", C.c_float), ("flip", C.c_float), ("mse_fixed", C.c_uint32),
                ("flip_fixed", C.c_uint32), ("ms_device", C.c_float)]


FLIP_DISABLED, FLIP_APPROX, FLIP_REFERENCE = 0, 1, 2

# every symbol include/vkgs_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "vkgs_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vkgs_destroy": (C.c_int, [C.c_void_p]),
    "vkgs_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vkgs_last_error": (C.c_char_p, [C.c_void_p]),
    "vkgs_version": (C.c_char_p, []),
    "vkgs_abi_struct_size": (C.c_uint32, [C.c_int]),
    "vkgs_pack_host": (C.c_int, [C.POINTER(SplatSetView), C.POINTER(Options), f32p, f32p, C.c_void_p, C.c_void_p]),
    "vkgs_upload": (C.c_int, [C.c_void_p, C.POINTER(SplatSetView), C.POINTER(Options)]),
    "vkgs_default_options": (None, [C.POINTER(Options)]),
    "vkgs_upload_scene": (C.c_int, [C.c_void_p, C.POINTER(SplatSetView), C.c_uint32, C.POINTER(Instance), C.c_uint32,
                                    C.POINTER(Options)]),
    "vkgs_set_instance_transform": (C.c_int, [C.c_void_p, C.c_uint32, f32p, f32p]),
    "vkgs_global_index_table": (C.c_int, [C.c_void_p, u32p, u32p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "vkgs_frame_params_from_camera": (C.c_int, [C.POINTER(Camera), C.c_uint32, C.c_uint32, C.POINTER(FrameParams)]),
    "vkgs_default_camera": (None, [C.POINTER(Camera)]),
    "vkgs_frame_params_set_fisheye": (None, [C.POINTER(FrameParams)]),
    "vkgs_quantize_normals_host": (C.c_int, [f32p, f32p, C.c_uint64]),
    "vkgs_exact_math_host": (C.c_int, [C.c_uint32, f32p, f32p, f32p, f32p, C.c_uint64]),
    "vkgs_render": (C.c_int, [C.c_void_p, C.POINTER(FrameParams), C.POINTER(Outputs)]),
    "vkgs_render_async": (C.c_int, [C.c_void_p, C.POINTER(FrameParams)]),
    "vkgs_render_to_host_async": (C.c_int, [C.c_void_p, C.POINTER(FrameParams), C.c_void_p]),
    "vkgs_set_frames_in_flight": (C.c_int, [C.c_void_p, C.c_int]),
    "vkgs_set_target_format": (C.c_int, [C.c_void_p, C.c_uint32]),
    "vkgs_sync": (C.c_int, [C.c_void_p]),
    "vkgs_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "vkgs_last_frame_stats": (C.c_int, [C.c_void_p, C.POINTER(Outputs)]),
    "vkgs_device_framebuffer": (C.c_void_p, [C.c_void_p]),
    "vkgs_launch_count": (C.c_uint64, [C.c_void_p]),
    "vkgs_sort_pairs": (C.c_int, [C.c_void_p, u32p, u32p, C.c_uint64, u32p, u32p, C.c_int, f32p]),
    "vkgs_sort_pairs_storage_bytes": (C.c_uint64, [C.c_uint64]),
    "vkgs_sort_pairs_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]),
    "vkgs_render_presorted": (C.c_int, [C.c_void_p, C.POINTER(FrameParams), u32p, C.c_uint64, C.POINTER(Outputs)]),
    "vkgs_capture_frame": (C.c_int, [C.c_void_p]),
    "vkgs_compare_with_capture": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(ImageMetrics)]),
    "vkgs_image_metrics_host": (C.c_int, [C.c_void_p, f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(ImageMetrics)]),
    "vkgs_read_surface_info": (C.c_int, [C.c_void_p, f32p, f32p, u32p]),
    "vkgs_read_records": (C.c_int, [C.c_void_p, u32p, C.c_uint64, C.c_uint64]),
    "vkgs_read_packed": (C.c_int, [C.c_void_p, f32p, f32p, f32p, f32p]),
    "vkgs_scene_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "vkgs_scene_view": (C.c_int, [C.c_void_p, C.POINTER(SplatSetView)]),
    "vkgs_scene_free": (C.c_int, [C.c_void_p]),
    "vkgs_scene_load_error": (C.c_char_p, []),
    "vkgs_synth_scene": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint64, f32p, f32p, f32p, f32p, f32p, f32p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libvkgs_b200.so (built in-tree by vk_gaussian_splatting_b200.build). Fails loudly."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m vk_gaussian_splatting_b200.build` "
                               "(the CUDA extension is required; there is no CPU fallback)")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib
