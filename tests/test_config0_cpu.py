"""BASELINE configs[0] — the CPU-only plumbing case: 100 k random Gaussians, SH degree 0, 512x512 pinhole, the reference's
CPU sorter (SplatSorterAsync::innerSort, src/splat_sorter_async.cpp:92-141) feeding a CPU blend in that order
(tryConsumeAndUploadCpuSortingResult semantics, src/splat_set_manager_vk.cpp:3334-3416: every splat drawn in the sorter's order,
frustum culling at the raster stage). No GPU: this is the oracle composed end to end, and the properties that tie it to the
GPU-key frame (dist.comp depth key + radix sort) of the same scene."""
import numpy as np

import vk_gaussian_splatting_b200 as g
from oracle import oracle as O


def _scene():
    s = g.synth_scene(100_000, 0, 0x3D650000)
    cam = g.default_camera()
    eye = np.array(cam.eye, np.float32)
    return s, cam, eye, np.array(cam.ctr, np.float32) - eye, np.eye(4, dtype=np.float32).reshape(16)


def test_config0_cpu_sorter_order_feeds_cpu_blend():
    s, cam, eye, direction, ident = _scene()
    pk = O.Packed(s)
    fp = O.frame_params(cam, 512, 512)
    for ftb in (0, 1):
        order, dist, ms_d, ms_s = O.cpu_sort(s.positions, ident, direction, eye, front_to_back=bool(ftb), mode=0, threads=1)
        # the sorter's contract: a permutation of all ids, monotone plane distance
        assert np.array_equal(np.sort(order), np.arange(s.size(), dtype=np.uint32))
        d = dist[order]
        assert np.all(d[1:] >= d[:-1]) if ftb else np.all(d[1:] <= d[:-1])
        img = O.render_presorted(pk, fp, O.default_options(front_to_back=ftb), order)
        assert np.isfinite(img).all() and img[..., 3].max() > 0.9
        # against the GPU-sorting mode of the same frame (NDC-depth key, dist-stage cull): the same picture — the two depth
        # measures only reorder splats whose depths nearly tie
        ref, keys, ids, _ = O.render(pk, fp, O.default_options(front_to_back=ftb))
        mse = float(np.mean((img[..., :3] - ref[..., :3]) ** 2))
        assert mse < 1e-4, mse
        if ftb:
            assert 0.0 <= img[..., 3].min() and img[..., 3].max() <= 1.0 + 1e-5
        # multi-threaded sorter (all host cores, __gnu_parallel::sort): same distances; the order may differ among exact ties only
        order_mt, dist_mt, _, _ = O.cpu_sort(s.positions, ident, direction, eye, front_to_back=bool(ftb), mode=1)
        assert np.array_equal(dist, dist_mt) and np.array_equal(dist[order], dist_mt[order_mt])


def test_presorted_oracle_equals_sorted_oracle_when_given_its_order():
    """Handing the oracle's own sorted ids to the presorted path reproduces the GPU-sorting frame bit for bit, as long as the
    dist-stage cull removed nothing the raster-stage cull would keep (default camera: it does not)."""
    s, cam, _, _, _ = _scene()
    pk = O.Packed(s)
    fp = O.frame_params(cam, 320, 240)
    for ftb in (0, 1):
        ref, keys, ids, _ = O.render(pk, fp, O.default_options(front_to_back=ftb))
        img = O.render_presorted(pk, fp, O.default_options(front_to_back=ftb), ids)
        assert np.array_equal(img, ref)


def test_rop16_rounding_stays_within_the_stated_bar():
    """Per-blend fp16 rounding (hardware ROP on the reference's default R16G16B16A16_SFLOAT target) against the fp32 target:
    within 4/255 per channel (SURVEY 8c iii) on a dense frame."""
    s = g.synth_scene(60_000, 3, 0x3D650018)
    cam = g.default_camera()
    pk = O.Packed(s)
    fp = O.frame_params(cam, 400, 225)
    for ftb in (0, 1):
        a, _, _, _ = O.render(pk, fp, O.default_options(front_to_back=ftb))
        b = O.render_rop16(pk, fp, O.default_options(front_to_back=ftb))
        d = np.abs(a[..., :3] - b[..., :3]).max()
        assert 0.0 < d <= 4.0 / 255.0
        assert np.array_equal(b, b.astype(np.float16).astype(np.float32))  # every texel is an fp16 value


def test_numpy_scene_restatement_matches_the_product_generator():
    """The CPU baselines generate their inputs without loading the product library: the numpy restatement of the
    synthetic-scene positions and the default camera must equal the product's bit for bit."""
    for n, seed in ((1000, 0x3D650001), (100_003, 0x3D650000)):
        assert np.array_equal(O.synth_positions(n, seed).view(np.uint32), g.synth_scene(n, 0, seed).positions.view(np.uint32))
    a, b = O.default_camera(), g.default_camera()
    assert bytes(a) == bytes(b)
