"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs. Bars: keys / ids / visible count / per-splat records bit-exact; RGBA within 1e-4
absolute per channel (fp32 target); full-size configs through size-independent properties."""
import ctypes as C

import numpy as np
import pytest

import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
from oracle import oracle as O
from util import f32_bits

pytestmark = pytest.mark.gpu

RGBA_TOL = 1e-4  # per-channel absolute tolerance, fp32 colour target


def _compare_frame(r, splats, cam, w, h, sh_degree=3, fp_set=None, tol=None, **optkw):
    opt = g.default_options(**optkw)
    r.upload(splats, opt)
    fp = g.frame_params(cam, w, h)
    fp.sh_degree = sh_degree
    for k, v in (fp_set or {}).items():
        setattr(fp, k, v)
    img, st, ids, keys = r.render(fp, want_sorted=True)
    rec = r.read_records()

    oopt = O.default_options(**{k: v for k, v in optkw.items() if k != "transmittance_epsilon"})  # the oracle never terminates early
    ofp = O.frame_params(cam, w, h)
    ofp.sh_degree = sh_degree
    for k, v in (fp_set or {}).items():
        setattr(ofp, k, v)
    pk = O.Packed(splats, sh_format=opt.sh_format, rgba_format=opt.rgba_format)
    oimg, okeys, oids, quads = O.render(pk, ofp, oopt, want_quads=True)

    # --- dist/cull + sort: bit-exact -----------------------------------------------------------
    assert st.visible_count == len(oids)
    assert np.array_equal(keys, okeys), "sorted keys differ"
    assert np.array_equal(ids, oids), "sort permutation differs"
    # --- per-splat records: bit-exact wherever the CUDA path kept the splat ------------------------
    vis = np.zeros(splats.size(), bool)
    vis[oids] = True
    bb0, bb1 = rec[:, 10], rec[:, 11]
    gvalid = vis & ((bb1 & 0xffff) >= (bb0 & 0xffff)) & ((bb1 >> 16) >= (bb0 >> 16))
    ovalid = vis & (quads["valid"] == 1)
    assert not np.any(gvalid & ~ovalid), "CUDA path rasterizes a splat the reference rejects"
    got = rec[gvalid, :10]
    want = np.concatenate([f32_bits(quads["center"]), f32_bits(quads["w1"]), f32_bits(quads["w2"]), f32_bits(quads["rgba"])], axis=1)[gvalid]
    assert np.array_equal(got, want), "per-splat records differ"
    # --- image ----------------------------------------------------------------------------------
    assert img.shape == oimg.shape
    diff = np.abs(img - oimg)
    if not optkw.get("front_to_back"):
        # BTF alpha is an unbounded sum of alphas: compare it relative to its magnitude
        diff[..., 3] /= np.maximum(1.0, np.abs(oimg[..., 3]))
    assert diff.max() <= (tol or RGBA_TOL), f"max abs diff {diff.max()}"
    return img, oimg, st


def test_config1_100k_deg0_512_btf(gpu_renderer):
    """BASELINE config[0]: 100k random Gaussians, SH degree 0, 512x512 (reference default: back-to-front)."""
    s = g.synth_scene(100_000, 0, 0x3D650000)
    img, oimg, st = _compare_frame(gpu_renderer, s, g.default_camera(), 512, 512, sh_degree=0)
    assert st.visible_count > 90_000 and img[..., 3].max() > 1.0


def test_config1_100k_deg0_512_ftb_exact(gpu_renderer):
    s = g.synth_scene(100_000, 0, 0x3D650000)
    _compare_frame(gpu_renderer, s, g.default_camera(), 512, 512, sh_degree=0, front_to_back=1)


def test_ftb_early_termination_within_stated_bound(gpu_renderer):
    """transmittance_epsilon = 2^-15: error <= eps * max|rgb| (rgb <= ~1.6 here) < 1e-4."""
    s = g.synth_scene(100_000, 3, 0x3D650001)
    cam = g.default_camera()
    r = gpu_renderer
    r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15))
    fp = g.frame_params(cam, 640, 360)
    img, st, _, _ = r.render(fp)
    pk = O.Packed(s)
    oimg, _, _, quads = O.render(pk, O.frame_params(cam, 640, 360), O.default_options(front_to_back=1), want_quads=True)
    cmax = np.abs(quads["rgba"][:, :3]).max()
    assert np.abs(img - oimg).max() <= max(RGBA_TOL, 2.0 ** -15 * cmax * 1.01)
    assert np.abs(img - oimg).max() <= RGBA_TOL


@pytest.mark.parametrize("w,h", [(640, 360), (333, 217), (16, 16), (1920, 1080)])
def test_deg3_scene_various_viewports(gpu_renderer, w, h):
    n = 60_000 if w < 1000 else 150_000
    s = g.synth_scene(n, 3, 0x3D650002)
    _compare_frame(gpu_renderer, s, g.default_camera(), w, h)


def test_orbit_cameras_and_model_transform(gpu_renderer):
    s = g.synth_scene(40_000, 3, 0x3D650004)
    for v in (1, 3, 6):
        _compare_frame(gpu_renderer, s, g.orbit_camera(v, 8), 480, 270)


@pytest.mark.parametrize("optkw", [
    dict(frustum_culling_mode=A.FRUSTUM_CULLING_NONE),
    dict(frustum_culling_mode=A.FRUSTUM_CULLING_AT_RASTER),
    dict(ms_antialiasing=1),
    dict(point_cloud_mode=1),
    dict(show_sh_only=1),
    dict(disable_opacity_gaussian=1, front_to_back=1),
    dict(sh_format=A.FORMAT_FLOAT16, rgba_format=A.FORMAT_FLOAT16),
    dict(sh_format=A.FORMAT_UINT8, rgba_format=A.FORMAT_UINT8),
    dict(size_culling_mode=A.SIZE_CULLING_ENABLED),
])
def test_shader_macro_variants(gpu_renderer, optkw):
    s = g.synth_scene(30_000, 3, 0x3D650005)
    cam = g.make_camera((0.4, 0.3, 1.6), (0, 0, 0)) if "frustum_culling_mode" in optkw else g.default_camera()
    _compare_frame(gpu_renderer, s, cam, 400, 300, **optkw)


def test_size_culling_threshold_straddle(gpu_renderer):
    """SIZE_CULLING_MODE (dist.comp.slang:93-134) with splats on both sides of sizeCullingMinPixels: the projected size
    goes through exp(scale); kernel and oracle run the same fixed-sequence exp, so V, keys and ids stay bit-exact even for
    splats within an ulp of the threshold. Thresholds are chosen so that a large share of the scene is culled."""
    s = g.synth_scene(120_000, 3, 0x3D650015)
    cam = g.default_camera()
    seen = []
    for min_px in (1.0, 6.0, 14.5, 40.0):
        _, _, st = _compare_frame(gpu_renderer, s, cam, 640, 360, fp_set=dict(size_culling_min_pixels=min_px),
                                  size_culling_mode=A.SIZE_CULLING_ENABLED)
        seen.append(st.visible_count)
    assert seen[0] > seen[1] > seen[2] > seen[3] > 0 and seen[3] < seen[0] // 2
    # a splat set whose projected sizes sit exactly AT the threshold: every splat at the same view depth, scales
    # spread over a few ulps around the value where projectedPixels == minPixels
    n = 4096
    t = g.synth_scene(n, 0, 0x3D650016)
    t.positions[:, 2] = 0.0
    t.positions[:, :2] *= 0.05
    camz = g.make_camera((0, 0, 3), (0, 0, 0))
    fpz = g.frame_params(camz, 256, 256)
    focal = max(abs(fpz.focal[0]), abs(fpz.focal[1]))
    crit = np.float32(np.log(2.0 * 3.0 / (2.8284271247 * 2.0 * focal)))  # extent * focal / dist == 2 px
    steps = np.arange(n, dtype=np.int32) - n // 2
    t.scale[:] = (np.full(n, crit, np.float32).view(np.int32) + steps // 8).view(np.float32)[:, None]
    _, _, st = _compare_frame(gpu_renderer, t, camz, 256, 256, sh_degree=0, fp_set=dict(size_culling_min_pixels=2.0),
                              size_culling_mode=A.SIZE_CULLING_ENABLED)
    assert 0 < st.visible_count < n


@pytest.mark.parametrize("sh_degree", [0, 1, 2, 3])
def test_frame_sh_degree_clamp(gpu_renderer, sh_degree):
    s = g.synth_scene(20_000, 3, 0x3D650006)
    _compare_frame(gpu_renderer, s, g.default_camera(), 320, 240, sh_degree=sh_degree)


def test_camera_inside_scene_clipping_and_huge_splats(gpu_renderer):
    """Camera inside the cloud: splats behind the camera, closer than near (depth clip), and splats
    hundreds of pixels wide (2048-px clamp path)."""
    s = g.synth_scene(20_000, 3, 0x3D650007)
    s.scale[:50] = np.log(0.5)  # very large splats
    cam = g.make_camera((0.1, 0.05, 0.2), (0.0, 0.0, -1.0))
    _compare_frame(gpu_renderer, s, cam, 320, 200)
    _compare_frame(gpu_renderer, s, cam, 320, 200, front_to_back=1)


def test_tiny_and_ragged_inputs(gpu_renderer):
    for n in (1, 2, 31, 255, 256, 257, 4097):
        s = g.synth_scene(n, 3, 0x3D650008)
        _compare_frame(gpu_renderer, s, g.default_camera(), 96, 64)


def test_nothing_visible(gpu_renderer):
    s = g.synth_scene(1000, 0, 0x3D650009)
    cam = g.make_camera((0, 0, 50), (0, 0, 100))  # looking away
    r = gpu_renderer
    r.upload(s)
    img, st, ids, _ = r.render(g.frame_params(cam, 64, 64), want_sorted=True)
    assert st.visible_count == 0 and len(ids) == 0 and st.tile_pairs == 0
    assert not img.any()


def test_ties_are_broken_by_ascending_id(gpu_renderer):
    """Many splats at the same depth: the stable sort must keep ascending splat id among equal keys."""
    n = 5000
    s = g.synth_scene(n, 0, 0x3D65000A)
    s.positions[:, 2] = 0.0  # camera on the z axis: all splats at identical view depth
    cam = g.make_camera((0, 0, 3), (0, 0, 0))
    img, oimg, st = _compare_frame(gpu_renderer, s, cam, 256, 256, sh_degree=0)
    r = gpu_renderer
    _, _, ids, keys = r.render(g.frame_params(cam, 256, 256), want_sorted=True)
    same = keys[1:] == keys[:-1]
    assert same.sum() > n // 2
    assert np.all(ids[1:][same] > ids[:-1][same])


def test_tile_list_overflow_regrows(gpu_renderer):
    """Tile lists start at max(8N, 2^20) pairs; big splats blow through that and must regrow."""
    n = 3000
    s = g.synth_scene(n, 0, 0x3D65000B)
    s.scale[:] = np.log(0.6)
    s.opacity[:] = -3.0
    r = gpu_renderer
    r.upload(s, g.default_options(front_to_back=1))
    fp = g.frame_params(g.default_camera(), 1920, 1080)
    img, st, _, _ = r.render(fp)
    assert st.tile_pairs > (1 << 20)
    pk = O.Packed(s)
    oimg, _, _, _ = O.render(pk, O.frame_params(g.default_camera(), 1920, 1080), O.default_options(front_to_back=1))
    assert np.abs(img - oimg).max() <= RGBA_TOL


# ---- stand-alone radix sort (vrdx replacement) -------------------------------------------------------

@pytest.mark.parametrize("n", [0, 1, 2, 31, 4095, 4096, 4097, 8191, 8192, 8193, 16385, 100_003, 1_000_000])
def test_sort_pairs_bit_exact(gpu_renderer, n):
    rng = np.random.default_rng(n + 1)
    keys = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    vals = rng.integers(0, 1 << 32, size=n, dtype=np.uint64).astype(np.uint32)
    k, v, _ = gpu_renderer.sort_pairs(keys, vals)
    ok, ov = O.radix_sort_pairs(keys, vals)
    assert np.array_equal(k, ok) and np.array_equal(v, ov)


def test_sort_pairs_stability_with_heavy_ties(gpu_renderer):
    rng = np.random.default_rng(7)
    n = 300_000
    keys = rng.integers(0, 7, size=n).astype(np.uint32) * np.uint32(0x01010101)  # 7 distinct keys, all digits tied
    keys[::5] = 0xffffffff  # same bit pattern as the in-kernel padding
    vals = np.arange(n, dtype=np.uint32)
    k, v, _ = gpu_renderer.sort_pairs(keys, vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])


def test_sort_pairs_presorted_and_reversed(gpu_renderer):
    n = 50_000
    asc = np.arange(n, dtype=np.uint32) * np.uint32(83_000)
    for keys in (asc, asc[::-1].copy(), np.zeros(n, np.uint32)):
        k, v, _ = gpu_renderer.sort_pairs(keys, np.arange(n, dtype=np.uint32))
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k, keys[order]) and np.array_equal(v, order.astype(np.uint32))


# ---- full-size configurations through size-independent properties -------------------------------------

def _full_size_properties(r, n, w, h, seed, ftb):
    s = g.synth_scene(n, 3, seed)
    r.upload(s, g.default_options(front_to_back=ftb, transmittance_epsilon=2.0 ** -15 if ftb else 0.0))
    fp = g.frame_params(g.default_camera(), w, h)
    img, st, ids, keys = r.render(fp, want_sorted=True)
    pk = O.Packed(s)
    okeys, oids = O.dist_cull(pk, O.frame_params(g.default_camera(), w, h), O.default_options(front_to_back=ftb))
    assert st.visible_count == len(oids)
    # sortedness, permutation (checksum of ids, key multiset), stability
    assert np.all(keys[1:] >= keys[:-1])
    assert np.array_equal(np.sort(ids), oids)
    assert np.array_equal(np.sort(okeys), keys)
    same = keys[1:] == keys[:-1]
    assert np.all(ids[1:][same] > ids[:-1][same])
    sk, si = O.radix_sort_pairs(okeys, oids)
    assert np.array_equal(si, ids)
    assert np.isfinite(img).all()
    if ftb:
        assert img[..., 3].max() <= 1.0 + 1e-5 and img[..., 3].min() >= 0.0
    # idempotence: the same frame twice is bit-identical (deterministic pipeline, no atomics in ordering)
    img2, _, ids2, _ = r.render(fp, want_sorted=True)
    assert np.array_equal(ids, ids2) and np.array_equal(img, img2)
    return img, st


def test_config2_1m_deg3_1080p_properties(gpu_renderer):
    img, st = _full_size_properties(gpu_renderer, 1_000_000, 1920, 1080, 0x3D650001, ftb=1)
    assert st.visible_count > 900_000


def test_config2_1m_deg3_1080p_image_vs_oracle(gpu_renderer):
    """The metric configuration itself against the oracle (the oracle needs ~10 s for this one)."""
    s = g.synth_scene(1_000_000, 3, 0x3D650001)
    _compare_frame(gpu_renderer, s, g.default_camera(), 1920, 1080, front_to_back=1)


def test_bench_configuration_image_vs_oracle(gpu_renderer):
    """Exactly what bench.py times: 1 M splats, SH3, 1080p, front to back, transmittance_epsilon = 2^-15, four frames in
    flight through the asynchronous entry point, frame copied to pinned host memory — against the oracle's exact frame.
    Bound: eps * max|rgb| (SH colours reach ~1.6) stays under the 1e-4 colour tolerance."""
    import torch
    s = g.synth_scene(1_000_000, 3, 0x3D650001)
    cam, w, h = g.default_camera(), 1920, 1080
    r = gpu_renderer
    r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15))
    r.set_frames_in_flight(4)
    fp = g.frame_params(cam, w, h)
    bufs = [torch.empty((h, w, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(4)]
    for b in bufs:
        r.render_to_host_async(fp, b)
    r.sync()
    oimg, okeys, oids, quads = O.render(O.Packed(s), O.frame_params(cam, w, h), O.default_options(front_to_back=1), want_quads=True)
    cmax = float(np.abs(quads["rgba"][:, :3]).max())
    assert 2.0 ** -15 * cmax < RGBA_TOL
    for b in bufs:
        assert np.array_equal(b, bufs[0])  # the four slots produce the same bits
    assert np.abs(bufs[0] - oimg).max() <= RGBA_TOL
    img, st, ids, keys = r.render(fp, want_sorted=True)
    assert np.array_equal(img, bufs[0]) and np.array_equal(ids, oids) and np.array_equal(keys, okeys)


def test_config3_6m_deg3_4k_image_vs_oracle(gpu_renderer):
    """BASELINE configs[2]: 6 M splats, SH3, 3840x2160, front to back — the alpha-blend tolerance check itself: the whole 4K
    frame against the oracle (exact compositing), then the bench setting (eps 2^-15) against the same oracle frame."""
    s = g.synth_scene(6_000_000, 3, 0x3D650002)
    cam, w, h = g.default_camera(), 3840, 2160
    r = gpu_renderer
    img, oimg, st = _compare_frame(r, s, cam, w, h, front_to_back=1)
    assert st.visible_count > 5_000_000 and oimg[..., 3].max() > 0.99
    r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15))
    img_eps, _, _, _ = r.render(g.frame_params(cam, w, h))
    assert np.abs(img_eps - oimg).max() <= RGBA_TOL
    r.upload(g.synth_scene(1000, 0, 1), g.default_options())  # release the device buffers


def test_config3_6m_deg3_4k_properties(gpu_renderer):
    img, st = _full_size_properties(gpu_renderer, 6_000_000, 3840, 2160, 0x3D650002, ftb=1)
    assert st.visible_count > 5_000_000


def test_config5_30m_deg3_1080p_sort_properties(gpu_renderer):
    """BASELINE configs[4]: the 30 M-splat stress scene. Dist/cull + radix sort at full size through
    size-independent properties (sortedness, key multiset, permutation, tie order, idempotence); the oracle's
    dist stage runs on the raw positions (the packed centers ARE the positions, src/splat_set_vk.cpp:253-262)."""
    import types
    n = 30_000_000
    r = gpu_renderer
    s = g.synth_scene(n, 3, 0x3D650004)
    r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15))
    fp = g.frame_params(g.default_camera(), 1920, 1080)
    img, st, ids, keys = r.render(fp, want_sorted=True)
    light = types.SimpleNamespace(centers=s.positions, scale=s.scale, n=n)
    okeys, oids = O.dist_cull(light, O.frame_params(g.default_camera(), 1920, 1080), O.default_options(front_to_back=1))
    del s
    assert st.visible_count == len(oids) > 29_000_000
    assert np.all(keys[1:] >= keys[:-1])
    same = keys[1:] == keys[:-1]
    assert np.all(ids[1:][same] > ids[:-1][same])
    assert np.array_equal(np.sort(okeys), keys)
    # permutation: order-independent checksums of the id set (sum and xor) + exact equality against the oracle's stable sort
    assert int(ids.astype(np.uint64).sum()) == int(oids.astype(np.uint64).sum())
    assert int(np.bitwise_xor.reduce(ids)) == int(np.bitwise_xor.reduce(oids))
    sk, si = O.radix_sort_pairs(okeys, oids)
    assert np.array_equal(si, ids)
    assert np.isfinite(img).all() and 0.0 <= img[..., 3].min() and img[..., 3].max() <= 1.0 + 1e-5
    img2, _, ids2, _ = r.render(fp, want_sorted=True)
    assert np.array_equal(ids, ids2) and np.array_equal(img, img2)
    # release the 19 GB of device buffers for the tests that follow
    r.upload(g.synth_scene(1000, 0, 1), g.default_options())


# ---- frames in flight / asynchronous API -----------------------------------------------------------------

def test_frames_in_flight_match_sequential_frames(gpu_renderer):
    """Two frames overlapping on the two internal streams give bit-identical images to one-at-a-time."""
    import torch
    s = g.synth_scene(200_000, 3, 0x3D65000C)
    r = gpu_renderer
    r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15))
    cams = [g.orbit_camera(v, 8) for v in range(4)]
    fps = [g.frame_params(c, 800, 450) for c in cams]
    r.set_frames_in_flight(1)
    ref = [r.render(fp)[0].copy() for fp in fps]
    r.set_frames_in_flight(2)
    bufs = [torch.empty((450, 800, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in fps]
    for rep in range(3):
        for fp, b in zip(fps, bufs):
            r.render_to_host_async(fp, b)
        r.sync()
        for b, want in zip(bufs, ref):
            assert np.array_equal(b, want)
    # the synchronous call still works with two slots and reports per-frame stats
    img, st, ids, _ = r.render(fps[1], want_sorted=True)
    assert np.array_equal(img, ref[1]) and st.visible_count == len(ids)


def test_four_frames_in_flight_are_bit_identical_to_sequential(gpu_renderer):
    """The default of the asynchronous API (and of bench.py): four frames in flight, eight frames per sync (every slot is
    used twice), different cameras — each result equals the one-frame-at-a-time render bit for bit."""
    import torch
    s = g.synth_scene(300_000, 3, 0x3D65000F)
    r = gpu_renderer
    r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15))
    fps = [g.frame_params(g.orbit_camera(v, 8), 960, 540) for v in range(8)]
    r.set_frames_in_flight(1)
    ref = [r.render(fp)[0].copy() for fp in fps]
    r.set_frames_in_flight(4)
    bufs = [torch.empty((540, 960, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in fps]
    for rep in range(3):
        for b in bufs:
            b[:] = -1.0
        for fp, b in zip(fps, bufs):
            r.render_to_host_async(fp, b)
        r.sync()
        for k, (b, want) in enumerate(zip(bufs, ref)):
            assert np.array_equal(b, want), (rep, k)


def test_overflow_with_four_frames_in_flight_is_repaired_by_sync(gpu_renderer):
    """Tile lists overflow while four frames are in flight: vkgs_sync grows the lists and renders the affected frames again
    on their own slots, so every host buffer holds the complete frame when it returns. A slot that is reused before the
    sync (8 frames, 4 slots) is checked — and repaired the same way — at the moment it is reused (the enqueue waits for
    the frame that still lives there: the back-pressure of the asynchronous API), so no frame is lost either way."""
    import torch
    n = 3000
    s = g.synth_scene(n, 0, 0x3D65000B)
    s.scale[:] = np.log(0.6)
    s.opacity[:] = -3.0
    r = gpu_renderer
    cams = [g.orbit_camera(v, 8) for v in range(4)]
    w, h = 1920, 1080
    fps = [g.frame_params(c, w, h) for c in cams]
    pk = O.Packed(s)
    want = [O.render(pk, O.frame_params(c, w, h), O.default_options(front_to_back=1))[0] for c in cams]
    bufs = [torch.empty((h, w, 4), dtype=torch.float32, pin_memory=True).numpy() for _ in range(8)]
    # (a) one frame per slot between syncs: repaired transparently
    r.upload(s, g.default_options(front_to_back=1))  # fresh lists: max(8N, 2^20) pairs
    r.set_frames_in_flight(4)
    for fp, b in zip(fps, bufs):
        r.render_to_host_async(fp, b)
    r.sync()
    st = r.last_frame_stats()
    assert st.tile_pairs > (1 << 20)
    for b, o in zip(bufs, want):
        assert np.abs(b - o).max() <= RGBA_TOL
    # (b) slots reused before the sync: the overflowed frames are repaired when their slots are reused, the rest by the sync
    r.upload(s, g.default_options(front_to_back=1))
    r.set_frames_in_flight(4)
    for b in bufs:
        b[:] = -1.0
    for k, b in enumerate(bufs):
        r.render_to_host_async(fps[k % 4], b)
    r.sync()
    for k, b in enumerate(bufs):
        assert np.abs(b - want[k % 4]).max() <= RGBA_TOL


def test_caller_stream_sees_finished_frames(gpu_renderer):
    import torch
    s = g.synth_scene(50_000, 3, 0x3D65000D)
    stream = torch.cuda.Stream()
    r = g.GaussianSplatting(0, stream=stream.cuda_stream)
    r.upload(s, g.default_options(front_to_back=1))
    fp = g.frame_params(g.default_camera(), 320, 180)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(6):
        r.render_async(fp)
    e1.record(stream)
    e1.synchronize()  # the caller's stream waited for every frame
    assert e0.elapsed_time(e1) > 0.05
    st = r.last_frame_stats()
    assert st.visible_count > 40_000
    r.close()


def test_colour_target_formats(gpu_renderer):
    """RGBA16F (the reference's default COLOR_MAIN format) and RGBA8_UNORM targets: blended in fp32,
    rounded once. fp16: within one half-precision ulp of the oracle's fp32 frame; unorm8: within 1 LSB."""
    s = g.synth_scene(80_000, 3, 0x3D65000E)
    cam = g.default_camera()
    r = gpu_renderer
    r.upload(s, g.default_options(front_to_back=1))
    fp = g.frame_params(cam, 640, 360)
    ref32, _, _, _ = r.render(fp)
    oimg, _, _, _ = O.render(O.Packed(s), O.frame_params(cam, 640, 360), O.default_options(front_to_back=1))
    assert np.abs(ref32 - oimg).max() <= RGBA_TOL
    r.set_target_format(A.FORMAT_FLOAT16)
    h, _, _, _ = r.render(fp)
    assert h.dtype == np.float16 and np.array_equal(h, ref32.astype(np.float16))  # exactly the RN rounding of the fp32 frame
    want = oimg.astype(np.float16)
    ulp = np.abs(h.view(np.int16).astype(np.int32) - want.view(np.int16).astype(np.int32))
    assert ulp.max() <= 1
    r.set_target_format(A.FORMAT_UINT8)
    u, _, _, _ = r.render(fp)
    want8 = np.rint(np.clip(oimg, 0, 1) * 255.0)
    assert u.dtype == np.uint8 and np.abs(u.astype(np.int32) - want8).max() <= 1
    r.set_target_format(A.FORMAT_FLOAT32)
    again, _, _, _ = r.render(fp)
    assert np.array_equal(again, ref32)


def test_loaded_scene_files_render_like_the_oracle(gpu_renderer):
    """Loader -> pack -> render: a .ply (SH3), its .spz twin written by the reference's spz library,
    and a .splat file, each rendered through the C ABI and compared with the oracle on the same arrays."""
    from pathlib import Path
    gold = Path(__file__).resolve().parent / "golden"
    cam = g.make_camera((0.0, 0.5, 6.0), (0, 0, 0))
    for name in ("loader_le_shuffled.ply", "loader_v3_from_ref.spz", "loader_records.splat"):
        s = g.load_scene(gold / name)
        s.scale += np.float32(2.5)  # fixtures carry tiny random splats: enlarge them so they cover pixels
        img, oimg, st = _compare_frame(gpu_renderer, s, cam, 320, 240)
        assert st.visible_count > 50 and oimg[..., 3].max() > 0.05, name


def _trs(tx, ty, tz, yaw_deg, scale):
    """glm-style column-major transform (memory order m[col][row]): translate * rotateY * scale."""
    a = np.deg2rad(yaw_deg)
    rot = np.array([[np.cos(a), 0, np.sin(a), 0], [0, 1, 0, 0], [-np.sin(a), 0, np.cos(a), 0], [0, 0, 0, 1]], np.float64)
    m = np.eye(4)
    m[:3, 3] = (tx, ty, tz)
    m = m @ rot @ np.diag([scale, scale, scale, 1.0])
    return np.ascontiguousarray(m.T, np.float32), np.ascontiguousarray(np.linalg.inv(m).T, np.float32)


def test_multi_instance_scene_matches_oracle(gpu_renderer):
    """Several instances of two splat sets (different SH degrees, one set instanced twice): global ids,
    sort permutation and per-splat records bit-exact, image within tolerance; the global index table is
    the reference's (src/splat_set_manager_vk.cpp:2304-2360); moving an instance needs no re-upload."""
    r = gpu_renderer
    a, b = g.synth_scene(30_011, 3, 0x3D6500C1), g.synth_scene(7_777, 0, 0x3D6500C2)  # ragged sizes: tiles straddle nothing
    xf = [_trs(0, 0, 0, 0, 1.0), _trs(0.9, 0.1, -0.4, 35, 0.6), _trs(-0.8, -0.2, 0.5, -70, 0.8)]
    inst = [(0, *xf[0]), (1, *xf[1]), (0, *xf[2])]
    cam = g.default_camera()
    w, h = 640, 360
    for ftb in (0, 1):
        opt = g.default_options(front_to_back=ftb)
        r.upload_scene([a, b], inst, opt)
        gi, gl = r.global_index_table()
        assert np.array_equal(gi, np.repeat([0, 1, 2], [a.size(), b.size(), a.size()]))
        assert np.array_equal(gl, np.concatenate([np.arange(a.size()), np.arange(b.size()), np.arange(a.size())]))
        fp = g.frame_params(cam, w, h)
        img, st, ids, keys = r.render(fp, want_sorted=True)
        pk = [O.Packed(a), O.Packed(b)]
        oimg, okeys, oids = O.render_scene(pk, inst, O.frame_params(cam, w, h), O.default_options(front_to_back=ftb))
        assert st.visible_count == len(oids) > 30_000
        assert np.array_equal(keys, okeys) and np.array_equal(ids, oids)
        d = np.abs(img - oimg)
        if not ftb:
            d[..., 3] /= np.maximum(1.0, np.abs(oimg[..., 3]))
        assert d.max() <= RGBA_TOL
        # records of an instance == records of the same set rendered alone with that model matrix
        rec = r.read_records()
        q = O.project_splat(pk[1], 1234, _with_model(O.frame_params(cam, w, h), *xf[1]), O.default_options(front_to_back=ftb))
        if q.valid and (oids == a.size() + 1234).any():
            want = np.concatenate([f32_bits(np.array(q.center)), f32_bits(np.array(q.w1)), f32_bits(np.array(q.w2)), f32_bits(np.array(q.rgba))])
            assert np.array_equal(rec[a.size() + 1234, :10], want)
    # move instance 2; only the transform changes
    xf2 = _trs(-0.5, 0.3, 0.2, 10, 0.7)
    r.set_instance_transform(2, *xf2)
    img2, st2, ids2, keys2 = r.render(g.frame_params(cam, w, h), want_sorted=True)
    inst2 = [inst[0], inst[1], (0, *xf2)]
    oimg2, okeys2, oids2 = O.render_scene(pk, inst2, O.frame_params(cam, w, h), O.default_options(front_to_back=1))
    assert np.array_equal(ids2, oids2) and np.array_equal(keys2, okeys2) and np.abs(img2 - oimg2).max() <= RGBA_TOL
    # a single identity instance is the plain vkgs_upload path
    ident = np.eye(4, dtype=np.float32)
    r.upload_scene([a], [(0, ident, ident)], g.default_options(front_to_back=1))
    i1, s1, id1, k1 = r.render(g.frame_params(cam, w, h), want_sorted=True)
    r.upload(a, g.default_options(front_to_back=1))
    i0, s0, id0, k0 = r.render(g.frame_params(cam, w, h), want_sorted=True)
    assert np.array_equal(id0, id1) and np.array_equal(k0, k1) and np.array_equal(i0, i1)


def _with_model(fp, t, ti):
    fp.model[:] = np.asarray(t, np.float32).reshape(16).tolist()
    fp.model_inverse[:] = np.asarray(ti, np.float32).reshape(16).tolist()
    return fp


def test_image_metrics_match_oracle(gpu_renderer):
    """MSE accumulator bit-exact (integer adds of identically rounded per-pixel terms), PSNR identical,
    FLIP-approx within 1e-4 relative (libm vs device powf/expf differ in the last ulp); capture/compare
    on rendered frames (ImageCompare, src/image_compare.cpp:770-905)."""
    r = gpu_renderer
    rng = np.random.default_rng(11)
    for (h, w) in ((37, 53), (360, 640)):
        a = rng.random((h, w, 4), dtype=np.float32)
        b = np.clip(a + rng.normal(0, 0.05, a.shape).astype(np.float32), 0, 1)
        for mode in (A.FLIP_DISABLED, A.FLIP_APPROX):
            m = r.image_metrics(a, b, mode)
            mf, ff, mse, psnr, flip = O.image_metrics(a, b, mode)
            assert m.mse_fixed == mf and m.mse == mse and m.psnr == pytest.approx(psnr, abs=1e-5)
            assert abs(int(m.flip_fixed) - ff) <= max(4, 1e-4 * ff) and m.flip == pytest.approx(flip, rel=1e-4, abs=1e-6)
    # FLIP "reference" mode: multi-scale features; the widest window (131x131) only applies 65 px away from the border
    a = rng.random((150, 170, 4), dtype=np.float32)
    b = np.clip(a + rng.normal(0, 0.08, a.shape).astype(np.float32), 0, 1)
    m = r.image_metrics(a, b, A.FLIP_REFERENCE)
    mf, ff, mse, psnr, flip = O.image_metrics(a, b, A.FLIP_REFERENCE)
    assert m.mse_fixed == mf and abs(int(m.flip_fixed) - ff) <= max(8, 2e-4 * ff) and m.flip == pytest.approx(flip, rel=2e-4)
    assert m.flip != pytest.approx(O.image_metrics(a, b, A.FLIP_APPROX)[4], rel=1e-3)  # a different estimator
    # rendered frames: capture one camera, compare with a slightly different one
    s = g.synth_scene(40_000, 3, 0x3D6500D1)
    r.upload(s, g.default_options(front_to_back=1))
    cam0, cam1 = g.default_camera(), g.make_camera((1.72, 1.5, 1.68))
    img0, _, _, _ = r.render(g.frame_params(cam0, 480, 270))
    r.capture_frame()
    same = r.compare_with_capture(A.FLIP_APPROX)
    assert same.mse_fixed == 0 and same.flip_fixed == 0 and same.psnr == pytest.approx(99.99)
    img1, _, _, _ = r.render(g.frame_params(cam1, 480, 270))
    m = r.compare_with_capture(A.FLIP_APPROX)
    mf, ff, mse, psnr, flip = O.image_metrics(img0, img1, 1)
    assert m.mse_fixed == mf > 0 and m.psnr == pytest.approx(psnr, abs=1e-5) and m.flip == pytest.approx(flip, rel=1e-4)


def test_fragment_counters_are_consistent_and_do_not_change_the_frame(gpu_renderer):
    """Profiling variant of the blend kernel (options._reserved[0] & 128): same image bit for bit; every blended
    fragment belongs to an evaluated (list entry, 8x8 block) pair; without early termination the number of blended
    fragments equals the number of non-discarded fragments of the oracle's rasterizer (alpha channel of a BTF frame
    rendered with rgb = 0, a = fragment count is not available, so the oracle count comes from its quads)."""
    r = gpu_renderer
    s = g.synth_scene(20_000, 0, 0x3D6500E1)
    cam, w, h = g.default_camera(), 320, 180
    fp = g.frame_params(cam, w, h)
    r.upload(s, g.default_options(front_to_back=1))
    img0, st0, _, _ = r.render(fp)
    assert st0.fragments_blended == 0 and st0.list_entries_evaluated == 0
    opt = g.default_options(front_to_back=1)
    opt._reserved[0] = 128
    r.upload(s, opt)
    img1, st1, _, _ = r.render(fp)
    assert np.array_equal(img0, img1)
    assert 0 < st1.fragments_blended <= 64 * st1.list_entries_evaluated
    # oracle: count fragments passing both discards (threedgs_raster.frag.slang:242,258) splat by splat
    pk = O.Packed(s)
    ofp, oopt = O.frame_params(cam, w, h), O.default_options(front_to_back=1)
    _, _, oids, quads = O.render(pk, ofp, oopt, want_quads=True)
    ys, xs = np.mgrid[0:h, 0:w]
    px, py = xs + 0.5, ys + 0.5
    total = 0
    for i in oids:
        q = quads[i]
        if not q["valid"]:
            continue
        dx, dy = (px - q["center"][0]).astype(np.float32), (py - q["center"][1]).astype(np.float32)
        f1 = dx * q["w1"][0] + dy * q["w1"][1]
        f2 = dx * q["w2"][0] + dy * q["w2"][1]
        A_ = f1 * f1 + f2 * f2
        total += int(np.count_nonzero((A_ <= 8.0) & (np.exp(-0.5 * A_) * q["rgba"][3] > 1.0 / 255.0)))
    assert abs(st1.fragments_blended - total) <= max(8, 2e-4 * total)  # (numpy's rounding of A near the two thresholds)


def test_surface_info_side_outputs_match_oracle(gpu_renderer):
    """NEED_SURFACE_INFO (front to back): integrated normals, picked depth + transmittance and splat id per pixel
    against the oracle (threedgs_raster.mesh.slang:209-233, frag.slang:316-350); the colour frame is unchanged.
    Tolerances: normals / transmittance 1e-4 like the colour; picked depth exact where both pick the same fragment
    (a pick flips only when T lands within 1e-6 of the iso threshold); ids equal except at such near-threshold pixels."""
    r = gpu_renderer
    s = g.synth_scene(30_000, 3, 0x3D6500F1)
    # make some particles flat / needle-like so every branch of the normal computation runs
    s.scale[::97, 1] = np.log(1e-7)
    s.scale[::193, :2] = np.log(1e-7)
    cam, w, h = g.default_camera(), 400, 240
    fp = g.frame_params(cam, w, h)
    r.upload(s, g.default_options(front_to_back=1))
    img0, _, _, _ = r.render(fp)
    # (full-precision normals here; the reference's default 2x16-bit octahedral transport is the last test of this file)
    r.upload(s, g.default_options(front_to_back=1, surface_info=1, quantize_normals=0))
    img, st, ids, _ = r.render(fp, want_sorted=True)
    assert np.array_equal(img, img0)
    nrm, dt, sid = r.read_surface_info(w, h)
    pk = O.Packed(s)
    oimg, onrm, odt, osid, oids = O.render_surface(pk, s.rotation, O.frame_params(cam, w, h),
                                                   O.default_options(front_to_back=1, quantize_normals=0))
    assert np.array_equal(ids, oids)
    assert np.abs(img - oimg).max() <= RGBA_TOL and np.abs(nrm - onrm).max() <= RGBA_TOL
    assert np.abs(dt[..., 1] - odt[..., 1]).max() <= RGBA_TOL
    near_iso = np.abs(odt[..., 1] - 0.7) < 1e-5
    same_pick = (dt[..., 0] == odt[..., 0]) | near_iso
    assert same_pick.mean() > 0.9999 and (odt[..., 0] != 0).mean() > 0.3
    assert (sid == osid).mean() > 0.9999 and (osid != 0xffffffff).mean() > 0.3
    with pytest.raises(g.VkgsError):
        r.upload(s, g.default_options(front_to_back=0, surface_info=1))  # the reference has no BTF surface pass (non-stochastic)


def _compare_gut_frame(r, s, cam, w, h, **optkw):
    opt = g.default_options(pipeline=A.PIPELINE_3DGUT, **optkw)
    r.upload(s, opt)
    fisheye = optkw.get("camera_model") == A.CAMERA_FISHEYE
    fp = g.frame_params(cam, w, h, fisheye=fisheye)
    img, st, ids, keys = r.render(fp, want_sorted=True)
    pk = O.Packed(s)
    oimg, okeys, oids, quads = O.render_gut(pk, s.rotation, O.frame_params(cam, w, h, fisheye=fisheye), O.default_gut_options(**optkw),
                                            want_quads=True)
    assert st.visible_count == len(oids) and np.array_equal(ids, oids) and np.array_equal(keys, okeys)
    d = np.abs(img - oimg)
    if not optkw.get("front_to_back"):
        d[..., 3] /= np.maximum(1.0, np.abs(oimg[..., 3]))
    # The particle response is ill-conditioned in the length of the canonical ray origin |ro| = |camera - centre| / scale
    # (cross(rd, ro) cancels): two fp32 evaluation orders — the oracle's, the CUDA fast path's, or a Vulkan compiler's FMA
    # contraction of the reference shader — differ by about 2^-24 |ro| in alpha. Accept / reject decisions stay exact.
    ro_max = float((np.linalg.norm(s.positions - np.array(cam.eye, np.float32), axis=1) / np.exp(s.scale.min(axis=1))).max())
    tol = RGBA_TOL + 4e-8 * ro_max
    assert d.max() <= tol, f"max abs diff {d.max()} (tolerance {tol}, |ro| up to {ro_max:.0f})"
    return img, oimg, st, quads


def test_3dgut_pipeline_matches_oracle(gpu_renderer):
    """VK3DGUT raster path (unscented-transform projection + per-fragment ray / particle response, pinhole,
    EXTENT_CONIC): same dist + sort front end (ids / keys bit-exact), image within the colour tolerance, for
    both compositing orders, the mip-splatting option, other kernel degrees and a model transform."""
    r = gpu_renderer
    s = g.synth_scene(40_000, 3, 0x3D650101)
    cam = g.default_camera()
    img_b, _, st, quads = _compare_gut_frame(r, s, cam, 480, 270)
    assert st.visible_count > 39_000 and quads["valid"].sum() > 30_000
    img_f, _, _, _ = _compare_gut_frame(r, s, cam, 480, 270, front_to_back=1)
    assert np.abs(img_b[..., :3] - img_f[..., :3]).max() < 1e-3  # over == under for the colour
    _compare_gut_frame(r, s, cam, 333, 217, front_to_back=1, ms_antialiasing=1)
    for deg in (0, 1, 3, 4, 5, 8):
        _compare_gut_frame(r, g.synth_scene(8_000, 0, 0x3D650102), cam, 256, 144, front_to_back=1, kernel_degree=deg)
    _compare_gut_frame(r, s, g.orbit_camera(3, 8), 400, 300, front_to_back=1, disable_opacity_gaussian=1)
    # a larger frame: millions of fragments, so the guard bands around the two discard thresholds get exercised
    _compare_gut_frame(r, g.synth_scene(120_000, 3, 0x3D650103), g.orbit_camera(5, 8), 960, 540, front_to_back=1)
    # tiny splats: canonical ray origins tens of thousands of units long (the fast path's guard band scales with |ro|)
    tiny = g.synth_scene(30_000, 0, 0x3D650104)
    tiny.scale -= np.float32(3.0)
    _compare_gut_frame(r, tiny, cam, 320, 200, front_to_back=1)
    # not the same estimator as the 3DGS pipeline, but the same picture
    r.upload(s, g.default_options(front_to_back=1))
    img3, _, _, _ = r.render(g.frame_params(cam, 480, 270))
    mse = float(np.mean((img3[..., :3] - img_f[..., :3]) ** 2))
    assert 1e-7 < mse < 2e-3
    # EXTENT_EIGEN quads (centre +- b1 +- b2 from the eigen-decomposition of the projected covariance)
    _compare_gut_frame(r, s, cam, 480, 270, front_to_back=1, extent_projection=A.EXTENT_EIGEN)
    _compare_gut_frame(r, s, g.orbit_camera(2, 8), 333, 217, extent_projection=A.EXTENT_EIGEN, ms_antialiasing=1)
    # unsupported combinations fail loudly
    for kw in (dict(surface_info=1, front_to_back=1),):
        with pytest.raises(g.VkgsError):
            r.upload(s, g.default_options(pipeline=A.PIPELINE_3DGUT, **kw))


def test_3dgut_bench_configuration_image_vs_oracle(gpu_renderer):
    """The 3DGUT line of bench.py (`configs.cfg2_3dgut`): 1 M splats, SH3, 1080p, front to back, transmittance_epsilon
    2^-15, pinhole, EXTENT_CONIC, quadratic kernel — the fast path of the blend (thresholds in squared distance, guard band,
    exact re-evaluation) over a few hundred million fragments, four frames in flight, against the oracle's exact frame."""
    s = g.synth_scene(1_000_000, 3, 0x3D650001)
    cam, w, h = g.default_camera(), 1920, 1080
    r = gpu_renderer
    img, oimg, st, quads = _compare_gut_frame(r, s, cam, w, h, front_to_back=1, transmittance_epsilon=2.0 ** -15)
    assert st.visible_count > 990_000 and quads["valid"].sum() > 900_000
    r.set_frames_in_flight(4)
    fp = g.frame_params(cam, w, h)
    for _ in range(4):
        r.render_async(fp)
    r.sync()
    again, _, _, _ = r.render(fp)
    assert np.array_equal(again, img)


def test_3dgut_fisheye_camera_matches_oracle(gpu_renderer):
    """CAMERA_FISHEYE on the VK3DGUT path: fisheye dist-stage cull (ids / keys bit-exact), equidistant projection of the
    sigma points, generateFisheyeRay per pixel with the field-of-view discard; fixed-sequence atan2 / acos / sin / cos on
    both sides, so decisions are exact. Wide (170 degree) and narrow fields of view, both extents, both orders."""
    r = gpu_renderer
    s = g.synth_scene(40_000, 3, 0x3D650131)
    cam = g.default_camera()
    fe = dict(camera_model=A.CAMERA_FISHEYE)
    img, oimg, st, quads = _compare_gut_frame(r, s, cam, 480, 270, front_to_back=1, **fe)
    assert st.visible_count > 20_000 and quads["valid"].sum() > 15_000
    # outside the unit circle of normalised pixel coordinates every fragment is discarded
    yy, xx = np.mgrid[0:270, 0:480]
    u, v = (xx + 0.5) / 479.0 * 2 - 1, (yy + 0.5) / 269.0 * 2 - 1
    assert np.all(img[np.sqrt(u * u + v * v) > 1.001] == 0) and img[..., 3].max() > 0.5
    _compare_gut_frame(r, s, cam, 333, 217, **fe)
    wide = g.default_camera()
    wide.fov_deg = 170.0
    wide.eye[:] = (0.3, 0.2, 0.4)  # inside the cloud: splats all around, many beyond the maximum angle
    _compare_gut_frame(r, s, wide, 400, 300, front_to_back=1, **fe)
    _compare_gut_frame(r, s, g.orbit_camera(2, 8), 320, 320, front_to_back=1, extent_projection=A.EXTENT_EIGEN, ms_antialiasing=1, **fe)
    # fisheye on the 3DGS raster pipelines is rejected (they are pinhole only)
    with pytest.raises(g.VkgsError):
        r.upload(s, g.default_options(camera_model=A.CAMERA_FISHEYE))


def test_3dgut_multi_instance_scene_matches_oracle(gpu_renderer):
    """VK3DGUT with several splat-set instances: the fragment stage takes its rays into the model space of the
    entry's instance (threedgut_raster.frag.slang:112-121)."""
    r = gpu_renderer
    a, b = g.synth_scene(20_011, 3, 0x3D650111), g.synth_scene(6_007, 0, 0x3D650112)
    xf = [_trs(0, 0, 0, 0, 1.0), _trs(0.9, 0.1, -0.4, 35, 0.6), _trs(-0.8, -0.2, 0.5, -70, 0.8)]
    inst = [(0, *xf[0]), (1, *xf[1]), (0, *xf[2])]
    cam, w, h = g.default_camera(), 480, 270
    opt = g.default_options(pipeline=A.PIPELINE_3DGUT, front_to_back=1)
    r.upload_scene([a, b], inst, opt)
    img, st, ids, keys = r.render(g.frame_params(cam, w, h), want_sorted=True)
    oimg, okeys, oids = O.render_scene([O.Packed(a), O.Packed(b)], inst, O.frame_params(cam, w, h), O.default_gut_options(front_to_back=1),
                                       rotations=[a.rotation, b.rotation])
    assert np.array_equal(ids, oids) and np.array_equal(keys, okeys) and st.visible_count > 20_000
    ro_max = 6.0 / float(np.exp(min(a.scale.min(), b.scale.min())) * 0.6)
    assert np.abs(img - oimg).max() <= RGBA_TOL + 4e-8 * ro_max


def test_surface_info_quantized_normals_match_oracle(gpu_renderer):
    """QUANTIZE_NORMALS (the reference default, src/parameters.h:195): the per-splat normal reaches the fragment stage
    through its 2x16-bit octahedral code (shaders/octahedral_normal.h.slang). The kernel runs the function whose host
    instantiation tests/test_oracle_kat.py pins bit for bit against the oracle; a normal that differs in its last bit before
    quantisation (device expf of the log-scale) can land in the neighbouring 16-bit bucket (one step = 3.05e-5 in octahedral
    space, <= 5.3e-5 per component after decoding), hence the 1.5e-4 bar on the integrated normals."""
    r = gpu_renderer
    s = g.synth_scene(30_000, 3, 0x3D6500F2)
    s.scale[::89, 2] = np.log(1e-7)
    cam, w, h = g.orbit_camera(1, 8), 384, 256
    fp = g.frame_params(cam, w, h)
    opt = g.default_options(front_to_back=1, surface_info=1)
    assert opt.quantize_normals == 1
    r.upload(s, opt)
    img, st, ids, _ = r.render(fp, want_sorted=True)
    nrm, dt, sid = r.read_surface_info(w, h)
    pk = O.Packed(s)
    oimg, onrm, odt, osid, oids = O.render_surface(pk, s.rotation, O.frame_params(cam, w, h), O.default_options(front_to_back=1))
    _, onrm_full, _, _, _ = O.render_surface(pk, s.rotation, O.frame_params(cam, w, h), O.default_options(front_to_back=1, quantize_normals=0))
    assert np.array_equal(ids, oids) and np.abs(img - oimg).max() <= RGBA_TOL
    assert np.abs(nrm - onrm).max() <= 1.5e-4
    # most pixels are bit-close (a bucket flip touches about one splat in a thousand, i.e. about one pixel in a hundred);
    # and the quantisation itself is visible against the full-precision normals
    assert np.quantile(np.abs(nrm - onrm).max(axis=-1), 0.9) <= 2e-6
    assert 1e-6 < np.abs(onrm - onrm_full).max() < 3e-4
    assert (sid == osid).mean() > 0.9999


# ---- CPU-sorting mode: frames drawn in a caller-supplied order ----------------------------------------------------------

def test_presorted_order_matches_oracle_and_composes_config0(gpu_renderer):
    """vkgs_render_presorted (the reference's SORTING_CPU_ASYNC_MULTI consumption, src/splat_set_manager_vk.cpp:3334-3416):
    all N splats in the order the CPU sorter produced (SplatSorterAsync::innerSort restatement, plane distance, descending
    for back-to-front), frustum culling forced to the raster stage. BASELINE configs[0] composed on the GPU: 100 k
    Gaussians, SH0, 512x512, CPU sort order -> the same image as the oracle's CPU blend of that order."""
    r = gpu_renderer
    s = g.synth_scene(100_000, 0, 0x3D650000)
    cam, w, h = g.default_camera(), 512, 512
    eye = np.array(cam.eye, np.float32)
    ident = np.eye(4, dtype=np.float32).reshape(16)
    pk = O.Packed(s)
    for ftb in (0, 1):
        order, dist, _, _ = O.cpu_sort(s.positions, ident, np.array(cam.ctr, np.float32) - eye, eye, front_to_back=bool(ftb), mode=0, threads=1)
        r.upload(s, g.default_options(front_to_back=ftb))
        img, st = r.render_presorted(g.frame_params(cam, w, h), order)
        oimg = O.render_presorted(pk, O.frame_params(cam, w, h), O.default_options(front_to_back=ftb), order)
        assert st.visible_count == s.size()
        d = np.abs(img - oimg)
        if not ftb:
            d[..., 3] /= np.maximum(1.0, np.abs(oimg[..., 3]))
        assert d.max() <= RGBA_TOL
        # the GPU-sorted frame of the same scene is the same picture up to the different depth key (plane distance vs NDC z)
        img2, _, ids, _ = r.render(g.frame_params(cam, w, h), want_sorted=True)
        assert float(np.mean((img2[..., :3] - img[..., :3]) ** 2)) < 1e-4
    # a camera inside the scene: splats behind the camera are in the list and must be culled at the raster stage
    cam_in = g.make_camera((0.1, 0.05, 0.2), (0.0, 0.0, -1.0))
    eye = np.array(cam_in.eye, np.float32)
    s3 = g.synth_scene(30_000, 3, 0x3D650017)
    order, _, _, _ = O.cpu_sort(s3.positions, ident, np.array(cam_in.ctr, np.float32) - eye, eye, front_to_back=True, mode=0, threads=1)
    r.upload(s3, g.default_options(front_to_back=1))
    img, st = r.render_presorted(g.frame_params(cam_in, 400, 300), order)
    oimg = O.render_presorted(O.Packed(s3), O.frame_params(cam_in, 400, 300), O.default_options(front_to_back=1), order)
    assert np.abs(img - oimg).max() <= RGBA_TOL and oimg[..., 3].max() > 0.5
    # a partial list (the first 5000 ids only) and an empty one
    img, st = r.render_presorted(g.frame_params(cam_in, 400, 300), order[:5000])
    oimg = O.render_presorted(O.Packed(s3), O.frame_params(cam_in, 400, 300), O.default_options(front_to_back=1), order[:5000])
    assert st.visible_count == 5000 and np.abs(img - oimg).max() <= RGBA_TOL
    img, st = r.render_presorted(g.frame_params(cam_in, 400, 300), order[:0])
    assert st.visible_count == 0 and not img.any()
    with pytest.raises(g.VkgsError):
        r.render_presorted(g.frame_params(cam_in, 400, 300), np.array([30_000], np.uint32))  # id out of range


def test_sort_pairs_device_in_place_with_device_count(gpu_renderer):
    """vkgs_sort_pairs_device (vrdxCmdSortKeyValueIndirect, 3rdparty/vrdx/src/vk_radix_sort.cc:249-258): device buffers,
    count read on the device (smaller than max_count), in place, on the caller's stream; elements past the count untouched."""
    import torch
    r = gpu_renderer
    rng = np.random.default_rng(21)
    # (the last case: a device-side count beyond the host bound — the caller's bug — is clamped, nothing is read past the buffers)
    for max_count, count in ((1_000_000, 777_777), (8192, 8192), (5000, 1), (100_000, 0), (300_000, 300_000), (5000, 123_456)):
        keys = rng.integers(0, 1 << 32, size=max_count, dtype=np.uint64).astype(np.uint32)
        keys[: max_count // 3] &= np.uint32(0xff00ff)  # plenty of ties
        vals = np.arange(max_count, dtype=np.uint32)
        dk = torch.from_numpy(keys.view(np.int32)).cuda()
        dv = torch.from_numpy(vals.view(np.int32)).cuda()
        dc = torch.tensor([count], dtype=torch.int32).cuda()
        nbytes = r.sort_pairs_storage_bytes(max_count)
        storage = torch.full((nbytes,), 0xAB, dtype=torch.uint8, device="cuda")  # garbage on purpose
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        for rep in range(2):  # the second call sorts sorted data with the same (dirty) storage
            r.sort_pairs_device(dk.data_ptr(), dv.data_ptr(), dc.data_ptr(), max_count, storage.data_ptr(), nbytes, stream.cuda_stream)
        stream.synchronize()
        k = dk.cpu().numpy().view(np.uint32)
        v = dv.cpu().numpy().view(np.uint32)
        count = min(count, max_count)
        order = np.argsort(keys[:count], kind="stable")
        assert np.array_equal(k[:count], keys[:count][order]) and np.array_equal(v[:count], vals[:count][order])
        assert np.array_equal(k[count:], keys[count:]) and np.array_equal(v[count:], vals[count:])


def test_rgba16f_target_within_rop_rounding_bar(gpu_renderer):
    """the reference's default colour target is R16G16B16A16_SFLOAT and its ROP rounds to fp16 after EVERY
    blend; this path blends in fp32 and rounds once. Against an oracle render that rounds per blend, the RGBA16F frame stays
    within 4/255 per channel (measured far below), for both compositing orders."""
    r = gpu_renderer
    s = g.synth_scene(150_000, 3, 0x3D650018)
    cam, w, h = g.default_camera(), 800, 450
    pk = O.Packed(s)
    for ftb in (1, 0):
        r.upload(s, g.default_options(front_to_back=ftb, target_format=A.FORMAT_FLOAT16))
        img, _, _, _ = r.render(g.frame_params(cam, w, h))
        assert img.dtype == np.float16
        rop = O.render_rop16(pk, O.frame_params(cam, w, h), O.default_options(front_to_back=ftb))
        d = np.abs(img.astype(np.float32)[..., :3] - rop[..., :3])
        assert d.max() <= 4.0 / 255.0, d.max()
        if ftb:
            assert np.abs(img.astype(np.float32)[..., 3] - rop[..., 3]).max() <= 4.0 / 255.0


def test_synchronous_strip_copies_match_the_one_piece_frame(gpu_renderer):
    """A synchronous frame of 12 MB or more leaves the device in four strips of tile rows, each copied while the next is
    blended (context.cu::enqueueFrame); the asynchronous path blends and copies in one piece. Same bits either way, for
    every target format, both orders, a height that is not a whole number of tiles and pageable as well as pinned memory."""
    import torch
    r = gpu_renderer
    s = g.synth_scene(200_000, 3, 0x3D650021)
    cam = g.default_camera()
    for fmt, tdt, (w, h) in ((A.FORMAT_FLOAT32, torch.float32, (1600, 1001)), (A.FORMAT_FLOAT16, torch.float16, (1920, 1080)),
                             (A.FORMAT_UINT8, torch.uint8, (2560, 1307))):
        for ftb in (1, 0):
            r.upload(s, g.default_options(front_to_back=ftb, target_format=fmt))
            fp = g.frame_params(cam, w, h)
            assert w * h * 4 * (4, 2, 1)[fmt] >= 12 << 20
            whole = torch.zeros((h, w, 4), dtype=tdt, pin_memory=True).numpy()
            r.set_frames_in_flight(2)
            r.render_to_host_async(fp, whole)
            r.sync()
            r.set_frames_in_flight(1)
            pinned = torch.zeros((h, w, 4), dtype=tdt, pin_memory=True).numpy()
            r.render(fp, out=pinned)
            pageable, _, _, _ = r.render(fp)
            assert whole.any() and np.array_equal(pinned, whole) and np.array_equal(pageable, whole)


def test_random_api_sequences_reproduce_fresh_context_frames(gpu_renderer):
    """Host-layer state machine: a seeded random walk over the entry points a caller mixes — synchronous frames, bursts of
    asynchronous frames (to device and to pinned host memory), viewport changes, frames-in-flight changes, colour-target
    changes, scene re-uploads with other options, caller-ordered frames — and after every step the frame must equal, bit
    for bit, what a second context that does nothing else renders for the same (scene, options, frame parameters)."""
    import torch
    rng = np.random.default_rng(0x3D650F00)
    scenes = [g.synth_scene(60_000, 3, 0x3D650F01), g.synth_scene(25_000, 0, 0x3D650F02)]
    optsets = [dict(front_to_back=1, transmittance_epsilon=2.0 ** -15), dict(front_to_back=0), dict(front_to_back=1, ms_antialiasing=1),
               dict(front_to_back=1, pipeline=A.PIPELINE_3DGUT), dict(front_to_back=1, sh_format=2, rgba_format=2)]
    sizes = [(640, 360), (333, 217), (1920, 1080), (1024, 1024), (64, 48)]
    cams = [g.default_camera(), g.orbit_camera(3, 8), g.orbit_camera(6, 8)]
    fmts = [(A.FORMAT_FLOAT32, torch.float32), (A.FORMAT_FLOAT16, torch.float16), (A.FORMAT_UINT8, torch.uint8)]
    r, fresh = gpu_renderer, g.GaussianSplatting(0)
    try:
        state = dict(scene=0, opt=0, size=0, cam=0, fmt=0, fif=1)
        expected = {}

        def upload(which):
            which.upload(scenes[state["scene"]], g.default_options(target_format=fmts[state["fmt"]][0], **optsets[state["opt"]]))

        def want():
            key = tuple(state[k] for k in ("scene", "opt", "size", "cam", "fmt"))
            if key not in expected:
                upload(fresh)
                w, h = sizes[state["size"]]
                expected[key] = fresh.render(g.frame_params(cams[state["cam"]], w, h))[0].copy()
            return expected[key]

        upload(r)
        r.set_frames_in_flight(1)
        for step in range(240):
            op = rng.integers(0, 9)
            w, h = sizes[state["size"]]
            fp = g.frame_params(cams[state["cam"]], w, h)
            if op == 0:
                state["size"] = int(rng.integers(0, len(sizes)))
            elif op == 1:
                state["cam"] = int(rng.integers(0, len(cams)))
            elif op == 2:
                state["fif"] = int(rng.integers(1, 5))
                r.set_frames_in_flight(state["fif"])
            elif op == 3:
                state["fmt"] = int(rng.integers(0, len(fmts)))
                r.set_target_format(fmts[state["fmt"]][0])
            elif op == 4:
                state["scene"], state["opt"] = int(rng.integers(0, len(scenes))), int(rng.integers(0, len(optsets)))
                upload(r)
                r.set_frames_in_flight(state["fif"])
            elif op == 5:
                img = r.render(fp)[0]
                assert np.array_equal(img, want()), f"step {step}: synchronous frame differs ({state})"
            elif op == 6:
                for _ in range(int(rng.integers(1, 7))):
                    r.render_async(fp)
                r.sync()
                img = r.render(fp)[0]
                assert np.array_equal(img, want()), f"step {step}: frame after an asynchronous burst differs ({state})"
            elif op == 7:
                bufs = [torch.zeros((h, w, 4), dtype=fmts[state["fmt"]][1], pin_memory=True).numpy() for _ in range(int(rng.integers(1, 6)))]
                for b in bufs:
                    r.render_to_host_async(fp, b)
                r.sync()
                for b in bufs:
                    assert np.array_equal(b, want()), f"step {step}: frame copied to pinned memory differs ({state})"
            elif op == 8 and optsets[state["opt"]].get("pipeline") != A.PIPELINE_3DGUT:
                # caller-ordered frame in the order of the last sorted frame: AT_RASTER culling, same picture within tolerance
                img, st, ids, _ = r.render(fp, want_sorted=True)
                pimg, _ = r.render_presorted(fp, ids)
                assert np.isfinite(np.asarray(pimg, np.float32)).all()
                assert np.array_equal(r.render(fp)[0], want()), f"step {step}: frame after a caller-ordered frame differs ({state})"
    finally:
        fresh.close()
        r.set_frames_in_flight(1)
        r.set_target_format(A.FORMAT_FLOAT32)


_SPECIAL = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-45, -1e-45, 1e-38, 3e38, -3e38, 1e20, -1e20, 1e-20, 88.0, -88.0, 104.0, -104.0], np.float32)


def _inject_special_values(s, attrs, rng, per_attr=40):
    for name in attrs:
        flat = getattr(s, name).reshape(-1)
        k = int(rng.integers(1, per_attr))
        flat[rng.integers(0, flat.size, k)] = _SPECIAL[rng.integers(0, _SPECIAL.size, k)]


@pytest.mark.parametrize("pipeline", [A.PIPELINE_3DGS, A.PIPELINE_3DGUT])
def test_special_values_in_geometry_match_oracle(gpu_renderer, pipeline):
    """NaN / +-inf / zero / denormal / huge values in positions, log-scales and rotations (a corrupt file): the visible count,
    keys and ids stay bit-exact and the image matches, non-finite in the same places. Rules both sides state explicitly: a
    NaN depth takes the canonical NaN of the reference's platform after the back-to-front negation (key 0xFFFFFFFF, sorted
    last in both orders); exp(NaN) is NaN; a 3DGUT particle so thin that the un-normalised fast path of the blend would
    overflow is evaluated in the oracle's operation order. (Non-finite colours / opacities are outside the contract: the
    reference's own result for them is undefined — SPIR-V min / max / comparison semantics for NaN; see the next test.)"""
    r = gpu_renderer
    w, h = 320, 200
    for t in range(12):
        rng = np.random.default_rng(0x3D650A00 + t)
        s = g.synth_scene(3000, 3, 0x3D650B00 + t)
        _inject_special_values(s, ("positions", "scale", "rotation"), rng)
        kw = dict(front_to_back=t & 1)
        if t % 3 == 0:
            kw["ms_antialiasing"] = 1
        if t % 5 == 0 and pipeline == A.PIPELINE_3DGS:
            kw["size_culling_mode"] = 1
        cam = g.orbit_camera(t % 8, 8)
        r.upload(s, g.default_options(pipeline=pipeline, **kw))
        img, st, ids, keys = r.render(g.frame_params(cam, w, h), want_sorted=True)
        if pipeline == A.PIPELINE_3DGUT:
            oimg, okeys, oids, _ = O.render_gut(O.Packed(s), s.rotation, O.frame_params(cam, w, h), O.default_gut_options(**kw))
        else:
            oimg, okeys, oids, _ = O.render(O.Packed(s), O.frame_params(cam, w, h), O.default_options(**kw))
        assert st.visible_count == len(oids) and np.array_equal(keys, okeys) and np.array_equal(ids, oids), f"trial {t}"
        fin = np.isfinite(oimg)
        assert np.array_equal(np.isfinite(img), fin), f"trial {t}: non-finite pixels in different places"
        d = np.abs(np.where(fin, img, 0) - np.where(fin, oimg, 0))
        if not kw["front_to_back"]:
            d[..., 3] /= np.maximum(1.0, np.abs(np.where(fin[..., 3], oimg[..., 3], 0)))
        # 3DGUT: the particle response is ill-conditioned in |ro| = distance / scale, unbounded for the thinnest particles here
        assert d.max() <= (RGBA_TOL if pipeline == A.PIPELINE_3DGS else 1e-3), f"trial {t}: max diff {d.max()}"


def test_non_finite_colours_stay_inside_their_splats_footprint(gpu_renderer):
    """Outside the parity contract but not outside the robustness one: splats with NaN / inf colours or opacities neither
    crash nor hang the frame, and every pixel outside their own pixel bounding boxes (where they are now, and where they
    were before they were corrupted: a NaN opacity makes a splat vanish) is bit-identical to the frame without them —
    discards are composited as zeros, so the per-splat kernel clamps non-finite colours to +-FLT_MAX. Both pipelines."""
    r = gpu_renderer
    cam, w, h = g.default_camera(), 640, 360
    fp = g.frame_params(cam, w, h)
    for pipeline in (A.PIPELINE_3DGS, A.PIPELINE_3DGUT):
        s = g.synth_scene(20_000, 3, 0x3D650C01)
        s.scale -= np.float32(1.0)  # small splats: most of the frame is not touched by the corrupted ones
        r.upload(s, g.default_options(front_to_back=1))
        r.render(fp)
        clean_rec = r.read_records()  # (pixel bounding boxes come from the 3DGS records; the 3DGUT quads are no larger)
        r.upload(s, g.default_options(front_to_back=1, pipeline=pipeline))
        clean, _, _, _ = r.render(fp)
        bad = np.random.default_rng(3).choice(s.size(), 12, replace=False)
        s.f_dc[bad[:4]] = np.float32(np.inf)
        s.f_rest[bad[4:8], 0] = np.float32(np.nan)
        s.opacity[bad[8:]] = np.float32(np.nan)
        r.upload(s, g.default_options(front_to_back=1))
        r.render(fp)
        rec = r.read_records()
        r.upload(s, g.default_options(front_to_back=1, pipeline=pipeline))
        img, st, _, _ = r.render(fp)
        mask = np.zeros((h, w), bool)
        for i in bad:
            for rr in (rec, clean_rec):
                bb0, bb1 = int(rr[i, 10]), int(rr[i, 11])
                x0, y0, x1, y1 = bb0 & 0xffff, bb0 >> 16, bb1 & 0xffff, bb1 >> 16
                if x1 >= x0 and y1 >= y0:
                    mask[max(0, y0 - 2):y1 + 3, max(0, x0 - 2):x1 + 3] = True
        assert mask.any() and mask.mean() < 0.2
        assert np.array_equal(img[~mask], clean[~mask]), f"pipeline {pipeline}"
        assert not np.array_equal(img[mask], clean[mask])


def test_random_options_cameras_and_viewports_match_oracle(gpu_renderer):
    """A slice of tools/fuzz_options.py (2050 trials were run clean on a B200 during development): random scene sizes
    (1 ... 20 000 splats), splat sizes, option combinations of both pipelines (culling modes, mip-splatting, storage formats,
    point cloud / SH-only / no-gaussian modes, kernel degrees, quad extents, fisheye), cameras (inside the cloud, far away,
    5 ... 170 degree fields of view, other near / far planes), viewports (1x1 ... 640x97) and per-frame parameters — ids /
    keys bit-exact and the image within tolerance in every trial."""
    failures = []
    assert _tool("fuzz_options").run(120, 0, gpu_renderer, log=lambda *a, **k: failures.append(" ".join(str(x) for x in a))) == 0, "\n".join(failures)


def _tool(name):
    import importlib.util
    import pathlib
    spec = importlib.util.spec_from_file_location(name, pathlib.Path(__file__).resolve().parent.parent / "tools" / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_random_multi_instance_scenes_match_oracle(gpu_renderer):
    """A slice of tools/fuzz_instances.py: 1-8 instances of 1-3 splat sets (ragged sizes, mixed SH degrees) under random
    rotations, translations, uniform / non-uniform / mirrored scales, random options and viewports."""
    failures = []
    assert _tool("fuzz_instances").run(40, 0, gpu_renderer, log=lambda *a, **k: failures.append(" ".join(str(x) for x in a))) == 0, "\n".join(failures)


def test_sort_pairs_fuzz_against_stable_argsort(gpu_renderer):
    """A slice of tools/fuzz_sort.py: sizes around the partition boundaries (0, 1, 4095 / 4096 / 4097, 65 536 +- 1, 2^20 + 1,
    3 000 001) and adversarial key distributions (constant, two values, sorted, reversed, few live bits, one live byte, one
    hot bin, a multiplicative hash), against numpy's stable argsort."""
    failures = []
    assert _tool("fuzz_sort").run(120, gpu_renderer, log=lambda *a, **k: failures.append(" ".join(str(x) for x in a))) == 0, "\n".join(failures)


def test_random_surface_info_frames_match_oracle(gpu_renderer):
    """A slice of tools/fuzz_surface.py: NEED_SURFACE_INFO side outputs (normals, picked depth + transmittance, splat id) on
    random scenes with flat and needle-like particles, cameras inside and outside the cloud, other iso / thin-particle
    thresholds, both normal transports."""
    failures = []
    assert _tool("fuzz_surface").run(40, 0, gpu_renderer, log=lambda *a, **k: failures.append(" ".join(str(x) for x in a))) == 0, "\n".join(failures)


def test_overflow_stress_with_state_changes_under_frames_in_flight(gpu_renderer):
    """A slice of tools/fuzz_overflow.py: bursts of asynchronous frames to pinned host memory, cameras alternating between few
    and very many tile pairs (every burst starts from the initial tile-list capacity, more frames than slots), ended by
    vkgs_sync or by one of the calls that change state under frames in flight (frames-in-flight count, colour target,
    scene upload: they complete the pending frames first, in the old state); every destination must hold exactly the
    frame a second context renders synchronously."""
    fresh = g.GaussianSplatting(0)
    try:
        failures = []
        assert _tool("fuzz_overflow").run(30, gpu_renderer, fresh, log=lambda *a, **k: failures.append(" ".join(str(x) for x in a))) == 0, "\n".join(failures)
    finally:
        fresh.close()
        gpu_renderer.set_target_format(A.FORMAT_FLOAT32)


def test_sort_pairs_repeated_uploads_are_ordered_before_the_sort(gpu_renderer):
    """Regression: vkgs_sort_pairs uploads keys and values from pageable memory on the legacy stream and sorts on a non-blocking
    stream; without a barrier in between the tail of the LAST upload (the values) could still be in flight when the sort
    read it — about one sort in ten at 300 k pairs on one box, keys always right, values corrupted. 200 sorts in a row."""
    r = gpu_renderer
    rng = np.random.default_rng(9)
    n = 294_912
    k = (np.arange(n, dtype=np.uint32) * np.uint32(2654435761)).astype(np.uint32)
    v = rng.permutation(n).astype(np.uint32)
    order = np.argsort(k, kind="stable")
    wk, wv = k[order], v[order]
    for it in range(200):
        ks, vs, _ = r.sort_pairs(k, v)
        assert np.array_equal(ks, wk) and np.array_equal(vs, wv), f"sort {it}"


def test_garbage_options_and_frame_parameters_are_rejected_or_rendered(gpu_renderer):
    """A slice of tools/fuzz_abi.py: invalid enums, NaN / inf / huge values and absurd sizes in vkgs_options and
    vkgs_frame_params — every call returns an error code or a frame (no crash, no hang), and afterwards the context still
    renders the reference frame bit for bit."""
    lines = []
    assert _tool("fuzz_abi").run(150, gpu_renderer, log=lambda *a, **k: lines.append(" ".join(str(x) for x in a))) == 0, "\n".join(lines)
