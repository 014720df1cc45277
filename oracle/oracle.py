"""ctypes/numpy wrapper of oracle/libvkgs_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
The struct layouts come from the product's public header through vk_gaussian_splatting_b200._abi
(plain-old-data only); every number is computed by oracle/vkgs_oracle.c and oracle/cpu_sorter.cpp.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from vk_gaussian_splatting_b200 import _abi as A

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libvkgs_oracle.so"

f32p, u32p = A.f32p, A.u32p


class Quad(C.Structure):
    _fields_ = [("center", C.c_float * 2), ("basis1", C.c_float * 2), ("basis2", C.c_float * 2), ("w1", C.c_float * 2),
                ("w2", C.c_float * 2), ("rgba", C.c_float * 4), ("ndc_z", C.c_float), ("valid", C.c_uint32)]


GUT_QUAD_DTYPE = np.dtype([("center", "<f4", 2), ("extent", "<f4", 2), ("basis1", "<f4", 2), ("basis2", "<f4", 2), ("rgba", "<f4", 4),
                           ("ndc_z", "<f4"), ("position", "<f4", 3), ("scale", "<f4", 3), ("inv_rot", "<f4", 9), ("valid", "<u4")])

QUAD_DTYPE = np.dtype([("center", "<f4", 2), ("basis1", "<f4", 2), ("basis2", "<f4", 2), ("w1", "<f4", 2), ("w2", "<f4", 2),
                       ("rgba", "<f4", 4), ("ndc_z", "<f4"), ("valid", "<u4")])

class OrcSet(C.Structure):
    _fields_ = [("centers", f32p), ("cov6", f32p), ("rgba", f32p), ("sh45", f32p), ("scale_log", f32p), ("n", C.c_uint64),
                ("sh_degree", C.c_uint32), ("_pad", C.c_uint32), ("rotation_wxyz", f32p)]


class OrcInstance(C.Structure):
    _fields_ = [("set_index", C.c_uint32), ("_pad", C.c_uint32), ("transform", C.c_float * 16),
                ("transform_inverse", C.c_float * 16)]


_lib = None


def build():
    subprocess.run(["make", "-C", str(HERE), "libvkgs_oracle.so"], check=True, stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        l = C.CDLL(str(LIB_PATH))
        l.orc_look_at.argtypes = [f32p, f32p, f32p, f32p]
        l.orc_perspective.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, f32p]
        l.orc_frame_params_from_camera.argtypes = [C.POINTER(A.Camera), C.c_uint32, C.c_uint32, C.POINTER(A.FrameParams)]
        l.orc_to_uint8.argtypes = [C.c_float, C.c_float, C.c_float]
        l.orc_to_uint8.restype = C.c_uint8
        l.orc_pack_half.argtypes = [C.c_float]
        l.orc_pack_half.restype = C.c_uint16
        l.orc_unpack_half.argtypes = [C.c_uint16]
        l.orc_unpack_half.restype = C.c_float
        l.orc_pack_cov6.argtypes = [f32p, f32p, f32p]
        l.orc_pack_rgba.argtypes = [f32p, C.c_float, f32p]
        l.orc_pack.argtypes = [C.POINTER(A.SplatSetView), C.c_uint32, C.c_uint32, f32p, f32p, f32p, f32p]
        l.orc_encode_min_max_fp32.argtypes = [C.c_float]
        l.orc_encode_min_max_fp32.restype = C.c_uint32
        l.orc_dist_cull.argtypes = [f32p, f32p, C.c_uint64, C.POINTER(A.FrameParams), C.POINTER(A.Options), u32p, u32p]
        l.orc_dist_cull.restype = C.c_uint32
        l.orc_radix_sort_pairs.argtypes = [u32p, u32p, C.c_uint64]
        l.orc_project_splat.argtypes = [C.c_uint32, f32p, f32p, f32p, f32p, C.c_uint32, C.POINTER(A.FrameParams),
                                        C.POINTER(A.Options), C.POINTER(Quad)]
        l.orc_expf.argtypes = [C.c_float]
        l.orc_expf.restype = C.c_float
        l.orc_atan2f_ypos.argtypes = [C.c_float, C.c_float]
        l.orc_atan2f_ypos.restype = C.c_float
        l.orc_acosf.argtypes = [C.c_float]
        l.orc_acosf.restype = C.c_float
        l.orc_sincosf.argtypes = [C.c_float, f32p, f32p]
        l.orc_raster_quad.argtypes = [C.POINTER(Quad), C.POINTER(A.FrameParams), C.POINTER(A.Options), f32p]
        l.orc_render.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_uint64, C.c_uint32, C.POINTER(A.FrameParams),
                                 C.POINTER(A.Options), f32p, u32p, u32p, C.c_void_p]
        l.orc_render.restype = C.c_uint32
        l.orc_render_scene.argtypes = [C.POINTER(OrcSet), C.c_uint32, C.POINTER(OrcInstance), C.c_uint32,
                                       C.POINTER(A.FrameParams), C.POINTER(A.Options), f32p, u32p, u32p]
        l.orc_render_scene.restype = C.c_uint32
        l.orc_render_gut_scene.argtypes = l.orc_render_scene.argtypes
        l.orc_render_gut_scene.restype = C.c_uint32
        l.orc_render_gut.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_uint64, C.c_uint32, C.POINTER(A.FrameParams),
                                     C.POINTER(A.Options), f32p, u32p, u32p, C.c_void_p]
        l.orc_render_gut.restype = C.c_uint32
        l.orc_gut_quad_size.restype = C.c_uint32
        assert l.orc_gut_quad_size() == GUT_QUAD_DTYPE.itemsize
        l.orc_gut_fragment.argtypes = [C.c_void_p, C.c_float, C.c_float, C.POINTER(A.FrameParams), C.POINTER(A.Options), f32p]
        l.orc_gut_fragment.restype = C.c_int
        l.orc_splat_normal.argtypes = [C.c_uint32, f32p, f32p, f32p, C.POINTER(A.FrameParams), f32p]
        l.orc_render_surface.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p, C.c_uint64, C.c_uint32, C.POINTER(A.FrameParams),
                                         C.POINTER(A.Options), f32p, f32p, f32p, u32p, u32p]
        l.orc_render_surface.restype = C.c_uint32
        l.orc_image_metrics.argtypes = [f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
        l.orc_quad_size.restype = C.c_uint32
        l.orc_cpu_sort.argtypes = [f32p, C.c_uint64, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, u32p, f32p,
                                   C.POINTER(C.c_double), C.POINTER(C.c_double)]
        l.orc_cpu_sort.restype = C.c_int
        l.orc_hardware_concurrency.restype = C.c_int
        assert l.orc_quad_size() == C.sizeof(Quad) == QUAD_DTYPE.itemsize
        _lib = l
    return _lib


def _p(a):
    return a.ctypes.data_as(f32p) if a is not None else None


def _u(a):
    return a.ctypes.data_as(u32p) if a is not None else None


def frame_params(cam: A.Camera, w: int, h: int, fisheye: bool = False) -> A.FrameParams:
    fp = A.FrameParams()
    lib().orc_frame_params_from_camera(C.byref(cam), w, h, C.byref(fp))
    if fisheye:  # FISHEYE focal, src/gaussian_splatting.cpp:1239-1243
        fp.focal[0] = np.float32(1.0) * np.float32(fp.viewport[0]) / np.float32(fp.fov_rad)
        fp.focal[1] = np.float32(-1.0) * np.float32(fp.viewport[1]) / np.float32(fp.fov_rad)
    return fp


def default_options(**kw) -> A.Options:
    """Reference defaults restated independently of the product (src/parameters.h, shaderio.h)."""
    o = A.Options()
    o.frustum_culling_mode = A.FRUSTUM_CULLING_AT_DIST
    o.quantize_normals = 1  # prmRaster.quantizeNormals, src/parameters.h:195
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Packed:
    """Device-layout arrays of a splat set as the reference's initDataBuffers would build them,
    decoded to fp32 the way the shaders fetch them."""

    def __init__(self, splats, sh_format=A.FORMAT_FLOAT32, rgba_format=A.FORMAT_FLOAT32):
        n = splats.size()
        self.n = n
        self.sh_degree = 3 if splats.f_rest.shape[1] == 45 else 0
        self.centers = np.empty((n, 3), np.float32)
        self.cov6 = np.empty((n, 6), np.float32)
        self.rgba = np.empty((n, 4), np.float32)
        self.sh = np.empty((n, 45), np.float32) if self.sh_degree else None
        self.scale = np.ascontiguousarray(splats.scale, np.float32)
        v = splats.view()
        lib().orc_pack(C.byref(v), sh_format, rgba_format, _p(self.centers), _p(self.cov6), _p(self.rgba), _p(self.sh))


def dist_cull(packed: Packed, fp, opt):
    keys = np.empty(packed.n, np.uint32)
    ids = np.empty(packed.n, np.uint32)
    v = lib().orc_dist_cull(_p(packed.centers), _p(packed.scale), packed.n, C.byref(fp), C.byref(opt), _u(keys), _u(ids))
    return keys[:v].copy(), ids[:v].copy()


def radix_sort_pairs(keys, vals):
    k = np.ascontiguousarray(keys, np.uint32).copy()
    v = np.ascontiguousarray(vals, np.uint32).copy()
    lib().orc_radix_sort_pairs(_u(k), _u(v), k.size)
    return k, v


def render(packed: Packed, fp, opt, want_quads=False):
    """Full oracle frame. Returns (image [H,W,4], sorted_keys, sorted_ids, quads or None)."""
    img = np.zeros((fp.height, fp.width, 4), np.float32)
    keys = np.empty(packed.n, np.uint32)
    ids = np.empty(packed.n, np.uint32)
    quads = np.zeros(packed.n, QUAD_DTYPE) if want_quads else None
    v = lib().orc_render(_p(packed.centers), _p(packed.cov6), _p(packed.rgba), _p(packed.sh), _p(packed.scale), packed.n,
                         packed.sh_degree, C.byref(fp), C.byref(opt), _p(img), _u(keys), _u(ids),
                         quads.ctypes.data_as(C.c_void_p) if want_quads else None)
    return img, keys[:v].copy(), ids[:v].copy(), quads


def render_rop16(packed: Packed, fp, opt):
    """Oracle frame with the colour target rounded to fp16 after every blend (ROP on R16G16B16A16_SFLOAT)."""
    fn = lib().orc_render_rop16
    fn.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_uint64, C.c_uint32, C.POINTER(A.FrameParams), C.POINTER(A.Options), f32p]
    fn.restype = C.c_uint32
    img = np.zeros((fp.height, fp.width, 4), np.float32)
    fn(_p(packed.centers), _p(packed.cov6), _p(packed.rgba), _p(packed.sh), _p(packed.scale), packed.n, packed.sh_degree,
       C.byref(fp), C.byref(opt), _p(img))
    return img


def render_presorted(packed: Packed, fp, opt, ids):
    """The reference's CPU-sorting mode: draw `ids` (global splat ids) in the caller's order, cull at raster."""
    fn = lib().orc_render_presorted
    fn.argtypes = [f32p, f32p, f32p, f32p, C.c_uint64, C.c_uint32, C.POINTER(A.FrameParams), C.POINTER(A.Options), u32p, C.c_uint64, f32p]
    fn.restype = None
    order = np.ascontiguousarray(ids, np.uint32)
    img = np.zeros((fp.height, fp.width, 4), np.float32)
    fn(_p(packed.centers), _p(packed.cov6), _p(packed.rgba), _p(packed.sh), packed.n, packed.sh_degree, C.byref(fp), C.byref(opt),
       _u(order), order.size, _p(img))
    return img


def render_scene(packed_sets, instances, fp, opt, rotations=None):
    """Multi-instance oracle frame. `instances`: list of (set_index, transform[4,4], transform_inverse[4,4]) in glm
    column-major memory order. Returns (image, sorted_keys, sorted_global_ids). With `rotations` (one [N,4] wxyz array
    per set) and opt.pipeline == 3DGUT the VK3DGUT pipeline renders the scene."""
    sets = (OrcSet * len(packed_sets))()
    keep = [np.ascontiguousarray(r, np.float32) for r in rotations] if rotations is not None else None
    for i, p in enumerate(packed_sets):
        if keep is not None:
            sets[i].rotation_wxyz = _p(keep[i])
        sets[i].centers, sets[i].cov6, sets[i].rgba, sets[i].sh45, sets[i].scale_log = _p(p.centers), _p(p.cov6), _p(p.rgba), _p(p.sh), _p(p.scale)
        sets[i].n, sets[i].sh_degree = p.n, p.sh_degree
    inst = (OrcInstance * len(instances))()
    total = 0
    for k, (si, t, ti) in enumerate(instances):
        inst[k].set_index = int(si)
        inst[k].transform[:] = np.asarray(t, np.float32).reshape(16).tolist()
        inst[k].transform_inverse[:] = np.asarray(ti, np.float32).reshape(16).tolist()
        total += packed_sets[int(si)].n
    img = np.zeros((fp.height, fp.width, 4), np.float32)
    keys = np.empty(total, np.uint32)
    ids = np.empty(total, np.uint32)
    fn = lib().orc_render_gut_scene if (keep is not None and opt.pipeline == A.PIPELINE_3DGUT) else lib().orc_render_scene
    v = fn(sets, len(packed_sets), inst, len(instances), C.byref(fp), C.byref(opt), _p(img), _u(keys), _u(ids))
    return img, keys[:v].copy(), ids[:v].copy()


def oct_quantize_normal(n) -> np.ndarray:
    """QUANTIZE_NORMALS round trip of one unit normal (shaders/octahedral_normal.h.slang)."""
    v = np.ascontiguousarray(n, np.float32).copy()
    lib().orc_oct_quantize_normal(_p(v))
    return v


def render_threads() -> int:
    """Threads the whole-frame oracle renders use (row bands; ORC_THREADS overrides)."""
    lib().orc_threads.restype = C.c_int
    return int(lib().orc_threads())


def default_gut_options(**kw) -> A.Options:
    """Reference defaults of the 3DGUT pipeline restated independently (src/parameters.h:190,215)."""
    o = default_options(pipeline=A.PIPELINE_3DGUT, extent_projection=A.EXTENT_CONIC, kernel_degree=2)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def render_gut(packed: Packed, rotation, fp, opt, want_quads=False):
    """VK3DGUT oracle frame. Returns (image, sorted_keys, sorted_ids, quads or None)."""
    img = np.zeros((fp.height, fp.width, 4), np.float32)
    keys, ids = np.empty(packed.n, np.uint32), np.empty(packed.n, np.uint32)
    quads = np.zeros(packed.n, GUT_QUAD_DTYPE) if want_quads else None
    rot = np.ascontiguousarray(rotation, np.float32)
    v = lib().orc_render_gut(_p(packed.centers), _p(packed.rgba), _p(packed.sh), _p(packed.scale), _p(rot), packed.n, packed.sh_degree,
                             C.byref(fp), C.byref(opt), _p(img), _u(keys), _u(ids), quads.ctypes.data_as(C.c_void_p) if want_quads else None)
    return img, keys[:v].copy(), ids[:v].copy(), quads


def gut_fragment(quad_record, px, py, fp, opt):
    """(accepted, opacity) of one fragment of a 3DGUT quad (one element of the quads array)."""
    op = np.zeros(1, np.float32)
    rec = np.ascontiguousarray(quad_record)
    ok = lib().orc_gut_fragment(rec.ctypes.data_as(C.c_void_p), float(px), float(py), C.byref(fp), C.byref(opt), _p(op))
    return bool(ok), float(op[0])


def splat_normal(packed: Packed, rotation, idx: int, fp) -> np.ndarray:
    out = np.zeros(3, np.float32)
    rot = np.ascontiguousarray(rotation, np.float32)
    lib().orc_splat_normal(idx, _p(packed.centers), _p(packed.scale), _p(rot), C.byref(fp), _p(out))
    return out


def render_surface(packed: Packed, rotation, fp, opt):
    """Front-to-back frame with the surface-info side outputs.
    Returns (image [H,W,4], normals [H,W,4], depth_transmittance [H,W,2], splat_id [H,W], sorted_ids)."""
    h, w = fp.height, fp.width
    img, nrm = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    dt, sid = np.zeros((h, w, 2), np.float32), np.zeros((h, w), np.uint32)
    ids = np.empty(packed.n, np.uint32)
    rot = np.ascontiguousarray(rotation, np.float32)
    v = lib().orc_render_surface(_p(packed.centers), _p(packed.cov6), _p(packed.rgba), _p(packed.sh), _p(packed.scale), _p(rot),
                                 packed.n, packed.sh_degree, C.byref(fp), C.byref(opt), _p(img), _p(nrm), _p(dt), _u(sid), _u(ids))
    return img, nrm, dt, sid, ids[:v].copy()


def render_gut_surface(packed: Packed, rotation, fp, opt):
    """VK3DGUT front-to-back frame with the surface-info side outputs (groundwork: the CUDA path does not build this
    combination yet). Returns (image, normals [H,W,4], depth_transmittance [H,W,2], splat_id [H,W], sorted_ids)."""
    h, w = fp.height, fp.width
    img, nrm = np.zeros((h, w, 4), np.float32), np.zeros((h, w, 4), np.float32)
    dt, sid = np.zeros((h, w, 2), np.float32), np.zeros((h, w), np.uint32)
    ids = np.empty(packed.n, np.uint32)
    rot = np.ascontiguousarray(rotation, np.float32)
    fn = lib().orc_render_gut_surface
    fn.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_uint64, C.c_uint32, C.POINTER(A.FrameParams), C.POINTER(A.Options), f32p, f32p, f32p, u32p, u32p]
    fn.restype = C.c_uint32
    v = fn(_p(packed.centers), _p(packed.rgba), _p(packed.sh), _p(packed.scale), _p(rot), packed.n, packed.sh_degree, C.byref(fp),
           C.byref(opt), _p(img), _p(nrm), _p(dt), _u(sid), _u(ids))
    return img, nrm, dt, sid, ids[:v].copy()


def image_metrics(reference, current, flip_mode=0):
    """(mse_fixed, flip_fixed, mse, psnr, flip): the shader's accumulators + the read-back arithmetic of
    src/image_compare.cpp:874-905."""
    ref = np.ascontiguousarray(reference, np.float32)
    cur = np.ascontiguousarray(current, np.float32)
    h, w = ref.shape[:2]
    res = np.zeros(4, np.uint32)
    lib().orc_image_metrics(_p(ref), _p(cur), w, h, flip_mode, _u(res))
    mse = np.float32(res[0]) / np.float32(1000000000.0)
    psnr = np.float32(99.99) if mse < 1e-10 else min(np.float32(10.0) * np.log10(np.float32(1.0) / mse), np.float32(99.99))
    flip = np.float32((float(res[2]) / 1000000000.0) ** (1.0 / 3.0))
    return int(res[0]), int(res[2]), float(mse), float(psnr), float(flip)


def project_splat(packed: Packed, idx: int, fp, opt) -> Quad:
    q = Quad()
    lib().orc_project_splat(idx, _p(packed.centers), _p(packed.cov6), _p(packed.rgba), _p(packed.sh), packed.sh_degree,
                            C.byref(fp), C.byref(opt), C.byref(q))
    return q


def cpu_sort(positions, model, direction, cop, front_to_back=False, mode=1, threads=None):
    """Reference CPU sorter restatement. Returns (indices, distances, ms_dist, ms_sort)."""
    l = lib()
    pos = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
    n = pos.shape[0]
    threads = threads or l.orc_hardware_concurrency()
    idx = np.empty(n, np.uint32)
    dist = np.empty(n, np.float32)
    m = np.ascontiguousarray(model, np.float32).reshape(16)
    d = np.ascontiguousarray(direction, np.float32)
    c = np.ascontiguousarray(cop, np.float32)
    t0, t1 = C.c_double(0), C.c_double(0)
    rc = l.orc_cpu_sort(_p(pos), n, _p(m), _p(d), _p(c), int(front_to_back), mode, threads, _u(idx), _p(dist), C.byref(t0),
                        C.byref(t1))
    assert rc == 0
    return idx, dist, t0.value, t1.value


def hardware_concurrency() -> int:
    return lib().orc_hardware_concurrency()


# ---- synthetic scene of SURVEY.md 8(d), restated in numpy (so that the CPU baselines need nothing from the product) ----------

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _stream_bits(seed: int, stream: int, idx):
    with np.errstate(over="ignore"):
        base = _mix64(np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0xD1342543DE82EF95)))
        return _mix64(base + (np.asarray(idx, dtype=np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))


def synth_positions(n: int, seed: int) -> np.ndarray:
    """positions ~ U([-1,1]^3) of the deterministic synthetic scene (counter-based splitmix64, stream 1, 24-bit uniforms):
    bit-identical to the product's generator (tests/test_config0_cpu.py), computed here with numpy only."""
    u = (_stream_bits(seed, 1, np.arange(3 * n, dtype=np.uint64)) >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return (np.float32(2.0) * u - np.float32(1.0)).reshape(n, 3)


def default_camera() -> A.Camera:
    """The reference's default camera (src/camera_set.h:48-53): eye (1.7,1.5,1.7), centre 0, up +Y, vfov 60, near 0.1, far 2000."""
    cam = A.Camera()
    cam.eye[:] = [1.7, 1.5, 1.7]
    cam.ctr[:] = [0.0, 0.0, 0.0]
    cam.up[:] = [0.0, 1.0, 0.0]
    cam.fov_deg, cam.znear, cam.zfar = 60.0, 0.1, 2000.0
    return cam
