"""Known-answer tests of the CPU oracle itself: analytic cases that pin the
restated shader arithmetic independently of any GPU."""
import ctypes as C
import math

import numpy as np
import pytest

import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
from oracle import oracle as O
from util import f32_bits, ulp_diff


def enc(x):
    return O.lib().orc_encode_min_max_fp32(float(np.float32(x)))


def test_encode_min_max_fp32_bit_patterns():
    assert enc(0.0) == 0x80000000
    assert enc(-0.0) == 0x7FFFFFFF
    assert enc(1.0) == 0xBF800000
    assert enc(-1.0) == 0x407FFFFF
    assert enc(np.inf) == 0xFF800000
    assert enc(-np.inf) == 0x007FFFFF


def test_encode_min_max_fp32_is_monotonic():
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.normal(size=4000), rng.normal(size=1000) * 1e-30, rng.normal(size=1000) * 1e30,
                        [0.0, -0.0, 1.0, -1.0]]).astype(np.float32)
    v = np.unique(v)  # sorted ascending, unique values (-0.0 == 0.0 collapses)
    e = np.array([enc(x) for x in v], np.uint64)
    assert np.all(np.diff(e.astype(np.int64)) > 0)


def test_radix_sort_is_stable_ascending():
    rng = np.random.default_rng(2)
    keys = rng.integers(0, 1 << 32, size=50000, dtype=np.uint64).astype(np.uint32)
    keys[::3] = keys[0]  # many ties
    vals = np.arange(keys.size, dtype=np.uint32)
    k, v = O.radix_sort_pairs(keys, vals)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])


def test_expf_within_1ulp_of_libm_on_fragment_domain():
    x = np.linspace(-4.0, 0.0, 20001, dtype=np.float32)
    got = np.array([O.lib().orc_expf(float(v)) for v in x], np.float32)
    want = np.exp(x.astype(np.float64)).astype(np.float32)
    assert ulp_diff(got, want).max() <= 1
    assert O.lib().orc_expf(0.0) == 1.0


def _one_splat(pos, log_scale, quat=(1, 0, 0, 0), opacity=10.0, dc=(0.0, 0.0, 0.0), rest=None):
    rest = np.zeros((1, 0), np.float32) if rest is None else np.asarray(rest, np.float32).reshape(1, 45)
    return g.SplatSet(np.array([pos], np.float32), np.array([dc], np.float32), rest, np.array([opacity], np.float32),
                      np.array([[log_scale] * 3], np.float32), np.array([quat], np.float32))


def _front_camera(dist=4.0):
    return g.make_camera((0, 0, dist), (0, 0, 0), (0, 1, 0), 60.0, 0.1, 100.0)


def test_single_isotropic_splat_has_analytic_alpha_profile():
    W = H = 256
    sigma, dist = 0.05, 4.0
    s = _one_splat((0, 0, 0), math.log(sigma), opacity=20.0, dc=(1.0, -1.0, 0.25))
    fp = O.frame_params(_front_camera(dist), W, H)
    pk = O.Packed(s)
    q = O.project_splat(pk, 0, fp, O.default_options())
    assert q.valid == 1
    assert abs(q.center[0] - W / 2) < 1e-3 and abs(q.center[1] - H / 2) < 1e-3
    focal = H / (2 * math.tan(math.radians(30)))
    var_px = (sigma * focal / dist) ** 2 + 0.3  # EWA projection + 0.3 dilation (threedgs.h.slang:69-70)
    # isotropic: both eigenvalues equal up to sqrt(max(0.1, .)) quirk: lambda = var +- sqrt(0.1)
    l1, l2 = var_px + math.sqrt(0.1), var_px - math.sqrt(0.1)
    r1 = math.hypot(q.basis1[0], q.basis1[1])
    r2 = math.hypot(q.basis2[0], q.basis2[1])
    assert r1 == pytest.approx(math.sqrt(8 * l1), rel=1e-4)
    assert r2 == pytest.approx(math.sqrt(8 * l2), rel=1e-4)
    assert abs(q.basis1[0] * q.basis2[0] + q.basis1[1] * q.basis2[1]) < 1e-3  # orthogonal
    img, _, ids, _ = O.render(pk, fp, O.default_options())
    assert ids.tolist() == [0]
    a = float(pk.rgba[0, 3])
    rgb = pk.rgba[0, :3]
    assert rgb[0] == pytest.approx(0.5 + 0.28209479177387814, rel=1e-6) and rgb[1] == pytest.approx(0.5 - 0.28209479177387814, rel=1e-6)
    # along the first eigenvector the profile is exp(-d^2 / (2 lambda1))
    e1 = np.array([q.basis1[0], q.basis1[1]]) / r1
    e2 = np.array([q.basis2[0], q.basis2[1]]) / r2
    checked = 0
    for j in range(H):
        for i in range(W):
            d = np.array([i + 0.5 - q.center[0], j + 0.5 - q.center[1]])
            u, v = d @ e1, d @ e2
            A_ = u * u / l1 + v * v / l2
            alpha = math.exp(-0.5 * A_) * a
            if A_ < 7.99 and alpha > 1 / 255 * 1.001:
                assert img[j, i, 3] == pytest.approx(alpha, rel=2e-4)
                assert img[j, i, 0] == pytest.approx(alpha * rgb[0], rel=2e-4)  # over black, BTF
                checked += 1
            elif A_ > 8.01 or alpha < 1 / 255 * 0.999:
                assert img[j, i, 3] == 0.0
    assert checked > 150


def test_two_overlapping_splats_btf_equals_ftb_rgb():
    W = H = 128
    s1 = _one_splat((0.0, 0.0, 0.0), math.log(0.1), opacity=0.3, dc=(1.5, 0.0, -1.0))
    s2 = _one_splat((0.05, 0.02, 1.0), math.log(0.08), opacity=-0.2, dc=(-1.0, 1.2, 0.4))
    s = g.SplatSet(np.vstack([s1.positions, s2.positions]), np.vstack([s1.f_dc, s2.f_dc]), np.zeros((2, 0)),
                   np.concatenate([s1.opacity, s2.opacity]), np.vstack([s1.scale, s2.scale]), np.vstack([s1.rotation, s2.rotation]))
    fp = O.frame_params(_front_camera(4.0), W, H)
    pk = O.Packed(s)
    btf, _, ids_b, _ = O.render(pk, fp, O.default_options(front_to_back=0))
    ftb, _, ids_f, _ = O.render(pk, fp, O.default_options(front_to_back=1))
    assert ids_b.tolist() == [0, 1]  # farthest (z=0) first; camera at z=4 looks down -z
    assert ids_f.tolist() == [1, 0]
    assert np.abs(btf[..., :3] - ftb[..., :3]).max() < 1e-6
    both = (btf[..., 3] > 0) & (ftb[..., 3] > 0)
    assert both.sum() > 100
    # alpha differs by definition: BTF = a1 + a2, FTB = 1 - (1-a1)(1-a2)
    assert np.all(btf[..., 3] >= ftb[..., 3] - 1e-6)


def test_sh_degree1_axis_directions():
    """rgb += C1 * (-s0*y + s1*z - s2*x) with dir = normalize(center - cam) (storage.h.slang:122)."""
    C1 = 0.4886025119029199
    rest = np.zeros(45, np.float32)
    # channel-major source layout: R coefficients 0..14, G 15..29, B 30..44
    rest[0], rest[15 + 1], rest[30 + 2] = 0.5, 0.25, -0.75  # R: s0, G: s1, B: s2
    s = _one_splat((0, 0, 0), math.log(0.05), opacity=5.0, rest=rest)
    pk = O.Packed(s)
    assert pk.sh[0, 0] == 0.5 and pk.sh[0, 3 * 1 + 1] == 0.25 and pk.sh[0, 3 * 2 + 2] == -0.75  # coefficient-major, RGB inner
    for eye, d in (((0, 0, 4), (0, 0, -1)), ((4, 0, 0), (-1, 0, 0)), ((0, 4, 0.0001), (0, -1, 0))):
        fp = O.frame_params(g.make_camera(eye, (0, 0, 0), (0, 1, 0) if eye[1] == 0 else (0, 0, -1), 60, 0.1, 100), 64, 64)
        fp.sh_degree = 1
        q = O.project_splat(pk, 0, fp, O.default_options())
        x, y, z = d
        assert q.rgba[0] == pytest.approx(0.5 + C1 * (-0.5 * y), abs=1e-4)
        assert q.rgba[1] == pytest.approx(0.5 + C1 * (0.25 * z), abs=1e-4)
        assert q.rgba[2] == pytest.approx(0.5 + C1 * (0.75 * x), abs=1e-4)


def test_sh_degree3_matches_float64_evaluation():
    rng = np.random.default_rng(5)
    n = 200
    s = g.synth_scene(n, 3, 77)
    pk = O.Packed(s)
    fp = O.frame_params(g.default_camera(), 640, 360)
    C1 = 0.4886025119029199
    C2 = [1.0925484, -1.0925484, 0.3153916, -1.0925484, 0.5462742]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    cam = np.array(fp.camera_position, np.float64)
    for i in range(n):
        q = O.project_splat(pk, i, fp, O.default_options())
        if pk.rgba[i, 3] < 1 / 255:
            continue
        d = pk.centers[i].astype(np.float64) - cam
        x, y, z = d / np.linalg.norm(d)
        sh = pk.sh[i].astype(np.float64).reshape(15, 3)
        rgb = pk.rgba[i, :3].astype(np.float64) + C1 * (-sh[0] * y + sh[1] * z - sh[2] * x)
        rgb += C2[0] * x * y * sh[3] + C2[1] * y * z * sh[4] + C2[2] * (2 * z * z - x * x - y * y) * sh[5] + C2[3] * x * z * sh[6] + C2[4] * (x * x - y * y) * sh[7]
        rgb += (C3[0] * sh[8] * (3 * x * x - y * y) * y + C3[1] * sh[9] * x * y * z + C3[2] * sh[10] * (4 * z * z - x * x - y * y) * y
                + C3[3] * sh[11] * z * (2 * z * z - 3 * x * x - 3 * y * y) + C3[4] * sh[12] * x * (4 * z * z - x * x - y * y)
                + C3[5] * sh[13] * (x * x - y * y) * z + C3[6] * sh[14] * x * (x * x - 3 * y * y))
        assert np.allclose(np.array(q.rgba[:3]), rgb, atol=5e-6)


def test_dist_cull_frustum_and_depth_order():
    pts = np.array([[0, 0, 0], [0, 0, 1], [0, 0, 3.95], [0, 0, 5], [100, 0, 0], [0, 0, -2100]], np.float32)
    n = len(pts)
    s = g.SplatSet(pts, np.zeros((n, 3)), np.zeros((n, 0)), np.ones(n), np.full((n, 3), -3.0), np.tile([1, 0, 0, 0], (n, 1)))
    pk = O.Packed(s)
    fp = O.frame_params(_front_camera(4.0), 128, 128)
    keys, ids = O.dist_cull(pk, fp, O.default_options())
    # kept: 0,1 (in front), 2 (closer than near: ndc.z<0 but > -0.2? no: z=0.05 from camera -> ndc.z<-0.2 culled)
    assert 3 not in ids and 4 not in ids and 5 not in ids  # behind camera, off to the side, beyond far
    assert ids.tolist()[:2] == [0, 1]
    k0, k1 = keys[0], keys[1]
    assert k0 < k1  # BTF key = enc(-depth): farther splat (id 0) sorts first
    keys_f, _ = O.dist_cull(pk, fp, O.default_options(front_to_back=1))
    assert keys_f[0] > keys_f[1]
    keys_n, ids_n = O.dist_cull(pk, fp, O.default_options(frustum_culling_mode=A.FRUSTUM_CULLING_NONE))
    assert ids_n.tolist() == list(range(n))  # no culling at the dist stage


def test_nan_depth_takes_the_canonical_key_and_sorts_last_in_both_orders():
    """A NaN / inf position fails no cull comparison, so the splat is kept; its key is the encoding of the canonical quiet NaN
    0x7FFFFFFF of the reference's platform, taken AFTER the back-to-front negation: 0xFFFFFFFF in both orders (x86 would
    hand the oracle 0xFFC00000 or the operand's payload, and the two orders would disagree)."""
    pts = np.array([[0, 0, 0], [np.nan, 0, 0], [0, 0, 1], [0, np.inf, 0], [0, 0, -np.inf]], np.float32)
    n = len(pts)
    s = g.SplatSet(pts, np.zeros((n, 3)), np.zeros((n, 0)), np.ones(n), np.full((n, 3), -3.0), np.tile([1, 0, 0, 0], (n, 1)))
    pk = O.Packed(s)
    fp = O.frame_params(_front_camera(4.0), 128, 128)
    for ftb in (0, 1):
        keys, ids = O.dist_cull(pk, fp, O.default_options(front_to_back=ftb))
        by_id = dict(zip(ids.tolist(), keys.tolist()))
        assert by_id[1] == 0xFFFFFFFF and by_id[3] == 0xFFFFFFFF, by_id
        assert by_id[0] < 0xFFFFFFFF and by_id[2] < 0xFFFFFFFF


def test_depth_clip_and_alpha_cull_reject_quads():
    fp = O.frame_params(_front_camera(4.0), 128, 128)
    # alpha cull: sigmoid(-8) < 1/255
    s = _one_splat((0, 0, 0), math.log(0.05), opacity=-8.0)
    assert O.project_splat(O.Packed(s), 0, fp, O.default_options()).valid == 0
    # closer than the near plane but inside the dilated frustum -> kept by dist cull when dilation allows, clipped by depth
    s = _one_splat((0, 0, 3.92), math.log(0.01), opacity=5.0)
    pk = O.Packed(s)
    q = O.project_splat(pk, 0, fp, O.default_options())
    assert q.ndc_z < 0 and q.valid == 0


def test_cpu_sorter_orders_by_plane_distance():
    s = g.synth_scene(20000, 0, 9)
    cam = g.default_camera()
    eye = np.array(cam.eye, np.float32)
    d = np.array(cam.ctr, np.float32) - eye
    ident = np.eye(4, dtype=np.float32).reshape(16)
    for ftb in (False, True):
        for mode in (0, 1):
            idx, dist, _, _ = O.cpu_sort(s.positions, ident, d, eye, front_to_back=ftb, mode=mode, threads=4)
            assert sorted(idx.tolist()) == list(range(20000))
            ds = dist[idx]
            assert np.all(np.diff(ds) >= 0) if ftb else np.all(np.diff(ds) <= 0)
    want = np.abs((s.positions.astype(np.float64) - eye) @ (d / np.linalg.norm(d)))
    assert np.allclose(dist, want, atol=1e-5)


def test_multi_instance_oracle_reduces_to_single_instance_and_composes():
    """orc_render_scene (global index table + per-instance transform): one identity instance == orc_render;
    an instance with model M == orc_render with fp.model = M; ids of later instances are offset by the
    splat counts before them (src/splat_set_manager_vk.cpp:2304-2360)."""
    s = g.synth_scene(3000, 3, 77)
    pk = O.Packed(s)
    cam = g.default_camera()
    fp = O.frame_params(cam, 160, 90)
    opt = O.default_options(front_to_back=1)
    ident = np.eye(4, dtype=np.float32)
    img0, k0, i0, _ = O.render(pk, fp, opt)
    img1, k1, i1 = O.render_scene([pk], [(0, ident, ident)], fp, opt)
    assert np.array_equal(img0, img1) and np.array_equal(k0, k1) and np.array_equal(i0, i1)
    m = np.eye(4)
    m[:3, 3] = (0.3, -0.1, 0.2)
    t, ti = np.ascontiguousarray(m.T, np.float32), np.ascontiguousarray(np.linalg.inv(m).T, np.float32)
    fpm = O.frame_params(cam, 160, 90)
    fpm.model[:] = t.reshape(16).tolist()
    fpm.model_inverse[:] = ti.reshape(16).tolist()
    img2, k2, i2, _ = O.render(pk, fpm, opt)
    img3, k3, i3 = O.render_scene([pk], [(0, t, ti)], fp, opt)
    assert np.array_equal(img2, img3) and np.array_equal(i2, i3)
    # two instances: every id of the second is offset by the first's splat count, both present
    img4, k4, i4 = O.render_scene([pk], [(0, ident, ident), (0, t, ti)], fp, opt)
    assert np.array_equal(np.sort(i4[i4 < 3000]), np.sort(i0)) and np.array_equal(np.sort(i4[i4 >= 3000]) - 3000, np.sort(i2))
    assert np.all(k4[1:] >= k4[:-1])


def test_image_metrics_known_answers():
    """MSE / PSNR / FLIP-approx restatement (image_compare_metric.comp.slang, image_compare.cpp:874-905)."""
    rng = np.random.default_rng(3)
    a = rng.random((48, 64, 4), dtype=np.float32)
    # identical images: zero error, PSNR clamps to 99.99 dB, FLIP 0
    mf, ff, mse, psnr, flip = O.image_metrics(a, a, 1)
    assert (mf, ff, mse, flip) == (0, 0, 0.0, 0.0) and psnr == pytest.approx(99.99)
    # constant offset d on RGB: MSE = d^2 (alpha ignored), PSNR = -20 log10 d; truncation loses < 1 unit per pixel
    b = a.copy()
    b[..., :3] += np.float32(0.1)
    b[..., 3] = 0.0
    mf, ff, mse, psnr, flip = O.image_metrics(a, b, 0)
    n = 48 * 64
    assert 0.01 * 1e9 - n <= mf <= 0.01 * 1e9 + 40 and ff == 0
    assert mse == pytest.approx(0.01, rel=1e-3) and psnr == pytest.approx(20.0, abs=0.01)
    # float64 restatement of the per-pixel MSE sum
    c = rng.random((48, 64, 4), dtype=np.float32)
    mf, _, mse, _, _ = O.image_metrics(a, c, 0)
    want = np.sum((a[..., :3].astype(np.float64) - c[..., :3]) ** 2) / (n * 3)
    assert mse == pytest.approx(want, rel=1e-4)
    # FLIP approx, flat black vs flat white (no Sobel response), worked by hand in float64:
    # F_L = 0.2 k^(1/3) (1 - exp(-0.42 k^(1/3))), k = 5; YCxCz(white) = F_L (M, L-M, M-S) with the HPE row sums;
    # error = csf(1cpd) (|dY| + 0.4 |dCx| + 0.4 |dCz|); uniform image -> pooled value = per-pixel error
    z, o = np.zeros((32, 32, 4), np.float32), np.ones((32, 32, 4), np.float32)
    _, ff, _, _, flip = O.image_metrics(z, o, 1)
    kc = 5.0 ** (1 / 3)
    fl = 0.2 * kc * (1 - math.exp(-0.42 * kc))
    L, M, S = 0.31670331 + 0.70299344 - 0.01969366, 0.10938715 + 0.87060437 + 0.01990658, 0.01840087 + 0.10476914 + 0.87470614
    csf = math.exp(-0.5) / math.sqrt(1 + (1 / 4) ** 2)
    want = fl * csf * (M + 0.4 * abs(L - M) + 0.4 * abs(M - S))
    # (each of the 1024 pixels contributes uint(want^3 / 1024 * 1e9) = 1072 of 1072.8: truncation, as in the shader)
    assert ff == 1024 * int(np.float32(want) ** 3 / 1024 * 1e9) and flip == pytest.approx(want, rel=5e-4)
    # FLIP reference mode: on flat images every band-pass feature vanishes, so it equals the colour term alone
    assert O.image_metrics(z, o, 2)[4] == pytest.approx(want, rel=5e-4)
    # and a luminance step in one image only is seen by the multi-scale features (interior pixels)
    e = np.full((140, 140, 4), 0.25, np.float32)
    f = e.copy()
    f[:, 70:, :3] = 0.75
    assert O.image_metrics(e, f, 2)[4] > O.image_metrics(e, e, 2)[4] == 0.0
    # monotone in the size of the error
    f1 = O.image_metrics(a, np.clip(a + 0.02, 0, 1), 1)[4]
    f2 = O.image_metrics(a, np.clip(a + 0.10, 0, 1), 1)[4]
    assert 0 < f1 < f2 < 1


def test_surface_normal_known_answers():
    """computeEllipsoidNormalMaxDensityPlane restatement (threedgrt.h.slang:358-418): for an axis-aligned
    ellipsoid the normal is Sigma^-1 (camera - centre) normalised; a flat particle takes its thin axis,
    flipped towards the camera; a needle / point falls back to the view direction; unit length always."""
    cam = g.default_camera()
    fp = O.frame_params(cam, 64, 64)
    eye = np.array(cam.eye, np.float64)

    def one(center, scale, rot):
        s = g.SplatSet(np.array([center], np.float32), np.zeros((1, 3)), np.zeros((1, 0)), np.zeros(1), np.log(np.array([scale], np.float32)),
                       np.array([rot], np.float32))
        return O.splat_normal(O.Packed(s), s.rotation, 0, fp).astype(np.float64)

    c = np.array([0.1, -0.2, 0.3])
    n = one(c, (0.5, 0.1, 0.2), (1, 0, 0, 0))
    want = (eye - c) / np.array([0.5, 0.1, 0.2]) ** 2
    assert np.allclose(n, want / np.linalg.norm(want), atol=2e-6) and abs(np.linalg.norm(n) - 1) < 1e-6
    # rotation by 90 deg about z (w,x,y,z) = (cos45, 0, 0, sin45): the canonical x axis maps to +y
    q = (np.cos(np.pi / 4), 0, 0, np.sin(np.pi / 4))
    n = one(c, (0.5, 0.1, 0.2), q)
    R = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1]], np.float64)
    want = R @ ((R.T @ (eye - c)) / np.array([0.5, 0.1, 0.2]) ** 2)
    assert np.allclose(n, want / np.linalg.norm(want), atol=2e-6)
    # a non-normalised quaternion is normalised first
    assert np.allclose(one(c, (0.5, 0.1, 0.2), tuple(3 * x for x in q)), n, atol=2e-6)
    # flat particle (one scale below thinParticleThreshold = 1e-6): its thin axis, towards the camera
    n = one(c, (0.5, 1e-7, 0.2), (1, 0, 0, 0))
    assert np.allclose(n, [0, np.sign(eye[1] - c[1]), 0], atol=1e-6)
    # needle: minus the view direction
    n = one(c, (1e-7, 1e-7, 0.2), (1, 0, 0, 0))
    assert np.allclose(n, (eye - c) / np.linalg.norm(eye - c), atol=2e-6)


def test_surface_outputs_single_splat():
    """One opaque splat in front of the camera: inside its footprint the side outputs follow the fragment shader
    (frag.slang:316-350): normal.a == colour.a, transmittance == 1 - alpha, depth picked where T < 0.7, id set;
    outside: the clear values (0, (0,1), 0xffffffff)."""
    cam = g.default_camera()
    s = g.SplatSet(np.zeros((1, 3)), np.ones((1, 3)), np.zeros((1, 0)), np.array([6.0]), np.log(np.full((1, 3), 0.05, np.float32)),
                   np.array([[1, 0, 0, 0]], np.float32))
    fp = O.frame_params(cam, 96, 96)
    pk = O.Packed(s)
    img, nrm, dt, sid, ids = O.render_surface(pk, s.rotation, fp, O.default_options(front_to_back=1))
    hit = sid != 0xffffffff
    assert hit.any() and not hit.all() and set(np.unique(sid[hit])) == {0}
    assert np.array_equal(nrm[..., 3], img[..., 3]) and np.allclose(dt[..., 1], 1.0 - img[..., 3], atol=1e-6)
    assert np.all(dt[~hit] == [0.0, 1.0]) and np.all(nrm[~hit] == 0)
    q = O.project_splat(pk, 0, fp, O.default_options(front_to_back=1))
    picked = dt[..., 0] != 0
    assert picked.any() and np.all(dt[picked, 0] == np.float32(q.ndc_z)) and np.all(dt[picked, 1] < 0.7) and np.all(dt[hit & ~picked, 1] >= 0.7)
    n = O.oct_quantize_normal(O.splat_normal(pk, s.rotation, 0, fp))  # default: the normal travels as an octahedral code
    assert np.allclose(nrm[hit, :3], n[None, :] * img[hit, 3:4], atol=1e-6)


def test_3dgut_known_answers():
    """VK3DGUT restatement: (i) the unscented transform of an isotropic splat on the optical axis projects to the
    image centre with covariance (f s / z)^2 to first order; (ii) the particle response along the ray through the
    splat centre is 1 (clamped to alphaClamp * density), and exp(-d^2/2) at canonical distance d; (iii) kernel degrees."""
    cam = g.make_camera((0, 0, 3.0))
    fp = O.frame_params(cam, 400, 400)
    sc = 0.01
    s = g.SplatSet(np.zeros((1, 3)), np.zeros((1, 3)), np.zeros((1, 0)), np.array([8.0]), np.log(np.full((1, 3), sc, np.float32)),
                   np.array([[1, 0, 0, 0]], np.float32))
    pk = O.Packed(s)
    opt = O.default_gut_options(front_to_back=1)
    img, keys, ids, quads = O.render_gut(pk, s.rotation, fp, opt, want_quads=True)
    q = quads[0]
    assert q["valid"] == 1 and np.allclose(q["center"], [200, 200], atol=1e-3)
    f = abs(fp.focal[0])
    sigma_px = f * sc / 3.0
    # conic extent: min(3.33, sqrt(2 ln(a/0.01))) * sqrt(sigma^2 + 0.3)
    a = 1 / (1 + math.exp(-8.0))
    ef = min(3.33, math.sqrt(2 * math.log(a / 0.01)))
    assert np.allclose(q["extent"], ef * math.sqrt(sigma_px ** 2 + 0.3), rtol=2e-3)
    assert np.allclose(q["inv_rot"].reshape(3, 3), np.eye(3), atol=1e-7) and np.allclose(q["scale"], sc, rtol=1e-6)
    # the ray generator adds 0.5 to SV_Position.xy (restated as written): the ray through the splat centre belongs to
    # the fragment at (199.5, 199.5)
    ok, op = O.gut_fragment(q, 199.5, 199.5, fp, opt)
    assert ok and op == pytest.approx(min(0.99, a), rel=1e-6)
    # one pixel to the right: canonical distance = (3 / f) / sc * cos-ish -> exp(-d^2 / 2)
    ok, op1 = O.gut_fragment(q, 200.5, 199.5, fp, opt)
    d = (3.0 / f) / sc
    assert ok and op1 == pytest.approx(a * math.exp(-0.5 * d * d), rel=2e-3)
    # response below KERNEL_MIN_RESPONSE (0.0113) or alpha <= 1/255 is rejected
    far = 199.5 + math.sqrt(-2 * math.log(0.0113)) / d + 1.5
    assert not O.gut_fragment(q, far, 199.5, fp, opt)[0]
    # kernel degrees at the same fragment: generalized Gaussians of threedgrt.h.slang:83-127
    dist = d * d
    for deg, want in ((0, max(1 - 0.329630334487 * math.sqrt(dist), 0)), (1, math.exp(-1.5 * math.sqrt(dist))),
                      (3, math.exp(-0.166666666667 * dist * math.sqrt(dist))), (4, math.exp(-0.0555555555556 * dist * dist)),
                      (5, math.exp(-0.0185185185185 * dist * dist * math.sqrt(dist))), (8, math.exp(-0.000685871056241 * dist ** 4))):
        ok, o = O.gut_fragment(q, 200.5, 199.5, fp, O.default_gut_options(front_to_back=1, kernel_degree=deg))
        assert ok and o == pytest.approx(min(0.99, a * want), rel=3e-3), deg
    # the frame: alpha at the centre pixel = clamped opacity
    assert img[199, 199, 3] == pytest.approx(min(0.99, a), rel=1e-6)


def test_fixed_sequence_trigonometry_known_answers():
    """orc_atan2f_ypos / orc_acosf / orc_sincosf (the fisheye camera's elementary functions, fixed operation sequences the
    CUDA path reproduces bit for bit) against double-precision numpy: exact anchor values and <= 4e-7 absolute elsewhere."""
    l = O.lib()
    f32p = C.POINTER(C.c_float)
    s, c = np.zeros(1, np.float32), np.zeros(1, np.float32)

    def sincos(x):
        l.orc_sincosf(float(x), s.ctypes.data_as(f32p), c.ctypes.data_as(f32p))
        return float(s[0]), float(c[0])

    assert sincos(0.0) == (0.0, 1.0)
    assert l.orc_acosf(1.0) == 0.0 and l.orc_acosf(0.0) == float(np.float32(np.pi / 2))
    assert l.orc_atan2f_ypos(1.0, 0.0) == float(np.float32(np.pi / 2))
    assert abs(l.orc_atan2f_ypos(1.0, 1.0) - np.pi / 4) < 1e-7 and abs(l.orc_atan2f_ypos(1e-7, -1.0) - np.pi) < 3e-7
    xs = np.linspace(-1, 1, 4001, dtype=np.float32)
    assert max(abs(l.orc_acosf(float(x)) - np.arccos(np.float64(x))) for x in xs) < 4e-7
    for x in np.concatenate([np.linspace(-6.5, 6.5, 4001), np.linspace(-100, 100, 1001)]).astype(np.float32):
        sv, cv = sincos(x)
        assert abs(sv - np.sin(np.float64(x))) < 2e-7 and abs(cv - np.cos(np.float64(x))) < 2e-7
    rng = np.random.default_rng(5)
    for _ in range(4000):
        y, x = np.float32(10 ** rng.uniform(-7, 3)), np.float32(rng.choice([-1, 1]) * 10 ** rng.uniform(-7, 3))
        assert abs(l.orc_atan2f_ypos(float(y), float(x)) - np.arctan2(np.float64(y), np.float64(x))) < 4e-7


def test_3dgut_fisheye_known_answers():
    """CAMERA_FISHEYE of the VK3DGUT oracle: the fisheye focal (gaussian_splatting.cpp:1239-1243), a splat on the optical
    axis projects to the principal point, its response peaks at the pixel the projection names, pixels outside the unit
    circle of normalised coordinates are discarded, the dist stage culls beyond the maximum angle."""
    cam = g.default_camera()
    w, h = 160, 90
    fp = O.frame_params(cam, w, h, fisheye=True)
    assert fp.fov_rad == np.float32(math.radians(60.0))
    assert fp.focal[0] == np.float32(w) / np.float32(fp.fov_rad) and fp.focal[1] == -np.float32(h) / np.float32(fp.fov_rad)
    assert bytes(fp) == bytes(g.frame_params(cam, w, h, fisheye=True))  # host helper of the product == oracle restatement
    opt = O.default_gut_options(camera_model=A.CAMERA_FISHEYE, front_to_back=1)
    s = g.synth_scene(3000, 0, 11)
    s.positions[0] = (0.0, 0.0, 0.0)  # the camera looks at the origin: on the optical axis
    s.scale[0] = np.log(0.05)
    s.opacity[0] = 4.0
    img, keys, ids, quads = O.render_gut(O.Packed(s), s.rotation, fp, opt, want_quads=True)
    q0 = quads[0]
    assert q0["valid"] == 1 and abs(q0["center"][0] - w / 2) < 1e-3 and abs(q0["center"][1] - h / 2) < 1e-3
    ok, op = O.gut_fragment(q0, w / 2 + 0.5, h / 2 + 0.5, fp, opt)
    assert ok and op > 0.6  # (the ray convention pixel / (res - 1) puts the axis 0.7 sigma away from this pixel centre)
    # response peak of an off-axis splat lands next to its projected centre (half a pixel of quantisation + the up to
    # half a pixel by which the ray convention pixel / (res - 1) and the projection's focal = res / fov disagree)
    q = quads[quads["valid"] == 1]
    big = q[np.argmax(q["extent"][:, 0])]
    cx, cy = big["center"]
    best = max((O.gut_fragment(big, i + 0.5, j + 0.5, fp, opt)[1], i + 0.5, j + 0.5)
               for j in range(max(0, int(cy) - 5), min(h, int(cy) + 6)) for i in range(max(0, int(cx) - 5), min(w, int(cx) + 6))
               if O.gut_fragment(big, i + 0.5, j + 0.5, fp, opt)[0])
    assert abs(best[1] - cx) <= 1.5 and abs(best[2] - cy) <= 1.5
    # field-of-view discard: corners of the frame are outside the unit circle
    assert not O.gut_fragment(q0, 0.5, 0.5, fp, opt)[0]
    yy, xx = np.mgrid[0:h, 0:w]
    u, v = (xx + 0.5) / (w - 1) * 2 - 1, (yy + 0.5) / (h - 1) * 2 - 1
    assert np.all(img[np.sqrt(u * u + v * v) > 1.001] == 0) and img[..., 3].max() > 0.9
    # dist-stage cull: a camera inside the cloud keeps fewer splats than there are, and never one behind it
    inside = g.default_camera()
    inside.eye[:] = (0.3, 0.2, 0.4)
    fpi = O.frame_params(inside, w, h, fisheye=True)
    pk = O.Packed(s)
    keys_i, ids_i = O.dist_cull(pk, fpi, opt)
    assert 0 < len(ids_i) < len(s.positions)


def test_octahedral_normal_quantisation_known_answers_and_host_kernel_function():
    """QUANTIZE_NORMALS round trip (shaders/octahedral_normal.h.slang): axis normals survive exactly up to one 16-bit step,
    every result is a unit vector within the 2x16-bit resolution of its input, and the function the CUDA kernel runs
    (its host instantiation behind vkgs_quantize_normals_host) equals the oracle bit for bit."""
    for axis in np.eye(3, dtype=np.float32):
        for sign in (1.0, -1.0):
            q = O.oct_quantize_normal(sign * axis)
            assert np.abs(q - sign * axis).max() < 4e-5 and abs(np.linalg.norm(q) - 1) < 2e-7
    rng = np.random.default_rng(17)
    n = rng.normal(size=(50_000, 3)).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)
    n = np.concatenate([n, np.array([[0.70710678, 0.70710678, 0], [0.70710678, 0, -0.70710678], [-0.0, 0.6, -0.8], [0.6, -0.0, 0.8]], np.float32)])
    ref = np.stack([O.oct_quantize_normal(v) for v in n])
    assert np.abs(ref - n).max() < 1.3e-4 and np.abs(np.linalg.norm(ref, axis=1) - 1).max() < 3e-7
    # bottom hemisphere really goes through the wrap (sign of z preserved)
    assert np.all(np.sign(ref[np.abs(n[:, 2]) > 1e-3, 2]) == np.sign(n[np.abs(n[:, 2]) > 1e-3, 2]))
    out = np.empty_like(n)
    f32p = C.POINTER(C.c_float)
    assert A.lib().vkgs_quantize_normals_host(n.ctypes.data_as(f32p), out.ctypes.data_as(f32p), len(n)) == 0
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    # surface frame: quantisation changes the integrated normals by no more than the code's resolution
    s = g.synth_scene(3000, 0, 23)
    fp = O.frame_params(g.default_camera(), 160, 100)
    pk = O.Packed(s)
    _, nq, _, _, _ = O.render_surface(pk, s.rotation, fp, O.default_options(front_to_back=1))
    _, nf, _, _, _ = O.render_surface(pk, s.rotation, fp, O.default_options(front_to_back=1, quantize_normals=0))
    assert 0 < np.abs(nq - nf).max() < 3e-4


def test_kernel_exact_math_host_instantiations_equal_the_oracle_bitwise():
    """The fixed-sequence exp / atan2 / acos / sin / cos the kernels run at their discard and cull decisions are
    host+device functions (csrc/device_common.cuh: round-to-nearest intrinsics on the device, the same IEEE operators on
    the host); their host instantiations must equal the oracle's restatements bit for bit over the whole domain the
    path uses. (Together with the GPU parity tests this pins both halves: same source on host and device, same bits
    as the oracle on the host.)"""
    l, ol = A.lib(), O.lib()
    f32p = C.POINTER(C.c_float)
    rng = np.random.default_rng(29)

    def run(which, a, b=None):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32) if b is not None else None
        out, out2 = np.empty_like(a), np.empty_like(a)
        rc = l.vkgs_exact_math_host(which, a.ctypes.data_as(f32p), b.ctypes.data_as(f32p) if b is not None else None,
                                    out.ctypes.data_as(f32p), out2.ctypes.data_as(f32p), len(a))
        assert rc == 0
        return out, out2

    # exp: the fragment domain [-4, 0], the kernel-degree domains, and the clamped tails
    x = np.concatenate([rng.uniform(-4.5, 0.0, 100_000), rng.uniform(-90, 90, 50_000), [0.0, -0.0, -87.0, 88.0, -1e9, 1e9]]).astype(np.float32)
    got, _ = run(0, x)
    ref = np.array([ol.orc_expf(float(v)) for v in x], np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # exp(NaN) = NaN on both sides (the clamps would otherwise turn it into exp(-87)); +-inf take the clamped tails
    got, _ = run(0, np.array([np.nan, np.inf, -np.inf], np.float32))
    ol.orc_expf.restype = C.c_float
    assert np.isnan(got[0]) and np.isnan(ol.orc_expf(float("nan")))
    assert got[1] == np.float32(ol.orc_expf(float("inf"))) and got[2] == np.float32(ol.orc_expf(float("-inf"))) and got[2] > 0
    # atan2(y > 0, x)
    y = (10.0 ** rng.uniform(-7, 4, 100_000)).astype(np.float32)
    xx = (rng.choice([-1.0, 1.0], 100_000) * 10.0 ** rng.uniform(-7, 4, 100_000)).astype(np.float32)
    xx[:100] = 0.0
    got, _ = run(1, y, xx)
    ref = np.array([ol.orc_atan2f_ypos(float(a), float(b)) for a, b in zip(y, xx)], np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # acos on [-1, 1]
    c = np.concatenate([rng.uniform(-1, 1, 100_000), [-1.0, 1.0, 0.0, 0.5, -0.5]]).astype(np.float32)
    got, _ = run(2, c)
    ref = np.array([ol.orc_acosf(float(v)) for v in c], np.float32)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # sin / cos: angles of the fisheye ray (|phi| <= pi, theta <= pi) and beyond
    t = np.concatenate([rng.uniform(-np.pi, np.pi, 100_000), rng.uniform(-100, 100, 20_000), [0.0, -0.0]]).astype(np.float32)
    gs, gc = run(3, t)
    s_, c_ = np.zeros(1, np.float32), np.zeros(1, np.float32)
    rs, rc_ = np.empty_like(t), np.empty_like(t)
    for i, v in enumerate(t):
        ol.orc_sincosf(float(v), s_.ctypes.data_as(f32p), c_.ctypes.data_as(f32p))
        rs[i], rc_[i] = s_[0], c_[0]
    assert np.array_equal(gs.view(np.uint32), rs.view(np.uint32)) and np.array_equal(gc.view(np.uint32), rc_.view(np.uint32))
    assert l.vkgs_exact_math_host(7, t.ctypes.data_as(f32p), None, gs.ctypes.data_as(f32p), None, 1) == A.VKGS_ERR_INVALID_ARGUMENT


def test_3dgut_surface_outputs_known_answers():
    """Oracle groundwork for NEED_SURFACE_INFO on the VK3DGUT pipeline (threedgut_raster.frag.slang:129-133,194-222,
    particleProcessHitGutWithNormal): the colour frame equals the plain 3DGUT frame; normal.a == colour.a and T == 1 - alpha;
    the fragment normal of a regular or flat particle is the per-splat max-density-plane normal of the 3DGS surface pass
    (it depends on the ray origin only), a needle takes minus its own ray direction; depth is picked where T < 0.7."""
    cam = g.default_camera()
    w, h = 96, 96
    fp = O.frame_params(cam, w, h)
    opt = O.default_gut_options(front_to_back=1)

    def one(scale):
        s = g.SplatSet(np.zeros((1, 3)), np.ones((1, 3)), np.zeros((1, 0)), np.array([6.0]), np.log(np.array([scale], np.float32)),
                       np.array([[0.9, 0.1, -0.3, 0.2]], np.float32))
        pk = O.Packed(s)
        img0, _, _, _ = O.render_gut(pk, s.rotation, fp, opt)
        img, nrm, dt, sid, ids = O.render_gut_surface(pk, s.rotation, fp, opt)
        assert np.array_equal(img, img0)
        hit = sid != 0xffffffff
        assert hit.any() and not hit.all() and np.array_equal(nrm[..., 3], img[..., 3])
        assert np.allclose(dt[..., 1], 1.0 - img[..., 3], atol=1e-6) and np.all(dt[~hit] == [0.0, 1.0]) and np.all(nrm[~hit] == 0)
        picked = dt[..., 0] != 0
        assert np.all(dt[picked, 1] < 0.7) and np.all(dt[hit & ~picked, 1] >= 0.7)
        return s, pk, img, nrm, hit

    s, pk, img, nrm, hit = one((0.05, 0.02, 0.08))
    n = O.splat_normal(pk, s.rotation, 0, fp)  # (3DGS surface pass, libm exp of the scale: last-bit differences only)
    assert np.allclose(nrm[hit, :3], n[None, :] * img[hit, 3:4], atol=2e-6)
    s, pk, img, nrm, hit = one((0.05, 1e-7, 0.08))  # flat: its thin axis
    n = O.splat_normal(pk, s.rotation, 0, fp)
    assert np.allclose(nrm[hit, :3], n[None, :] * img[hit, 3:4], atol=2e-6)
    s, pk, img, nrm, hit = one((1e-7, 1e-7, 0.08))  # needle: minus the ray direction of each pixel -> varies over the footprint
    if hit.sum() > 1:
        dirs = nrm[hit, :3] / img[hit, 3:4]
        assert np.allclose(np.linalg.norm(dirs, axis=1), 1.0, atol=1e-5)
        to_cam = np.array(cam.eye, np.float32) / np.linalg.norm(cam.eye)
        assert np.all(dirs @ to_cam > 0.99)  # towards the camera, within the few degrees the footprint spans
