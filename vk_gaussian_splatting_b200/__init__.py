"""vk_gaussian_splatting_b200 — B200-native (sm_100a CUDA) 3DGS forward rasterization path.

Drop-in for the VK3DGSR render call of nvpro-samples/vk_gaussian_splatting (dist/cull -> radix sort
-> raster). The numeric path lives in lib/libvkgs_b200.so (C ABI: include/vkgs_b200.h); this package
is the thin host-side mirror of the reference's interface. No CPU fallback exists.
"""
from ._abi import (FORMAT_FLOAT16, FORMAT_FLOAT32, FORMAT_UINT8, FRUSTUM_CULLING_AT_DIST, FRUSTUM_CULLING_AT_RASTER,
                   FRUSTUM_CULLING_NONE, SIZE_CULLING_DISABLED, SIZE_CULLING_ENABLED, Camera, FrameParams, Options, Outputs,
                   SplatSetView, lib)
from .api import (FrameStats, GaussianSplatting, SplatSet, VkgsError, default_camera, default_options, frame_params,
                  load_scene, make_camera, orbit_camera, pack_host, synth_scene)

__all__ = ["GaussianSplatting", "SplatSet", "FrameStats", "VkgsError", "Camera", "FrameParams", "Options", "Outputs",
           "SplatSetView", "default_camera", "default_options", "frame_params", "make_camera", "orbit_camera",
           "synth_scene", "pack_host", "load_scene", "lib"]
