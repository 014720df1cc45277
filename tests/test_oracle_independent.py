"""Independent float64 cross-checks of the oracle (CPU only).

The C oracle restates the reference's shaders operation by operation in fp32. These tests re-derive the same frames a
second time, in vectorised float64 numpy and from the *mathematical* statement of each stage (SURVEY.md Appendix A for the
3DGS path; the ray / Gaussian Mahalanobis distance for the 3DGUT fragment) rather than from the C code, and require the two
to agree. A transcription slip in the oracle (a transposed matrix, a wrong sign, a swapped axis, a missing term) shows up
as a large difference; fp32-vs-fp64 rounding only flips a handful of fragments that sit within rounding of a discard
threshold. Neither side is the reference itself: the frame stays "parity unpinned" against a Vulkan device.
"""
import numpy as np

import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
from oracle import oracle as O

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484, -1.0925484, 0.3153916, -1.0925484, 0.5462742]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
      1.445305721320277, -0.5900435899266435]


def _mat(m16):
    """glm column-major float[16] -> ordinary 4x4 matrix acting on column vectors."""
    return np.array(m16, np.float64).reshape(4, 4).T


def _rotations(q_wxyz):
    q = q_wxyz.astype(np.float64)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                     np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], 1)  # [N,3,3]


def _colours(s, cam_model):
    """clamp(0.5 + C0 dc) + SH(dir), sigmoid(opacity) (Appendix A.1, A.3)."""
    n = s.size()
    rgb = np.clip(0.5 + C0 * s.f_dc.astype(np.float64), 0, 1)
    alpha = np.clip(1.0 / (1.0 + np.exp(-s.opacity.astype(np.float64))), 0, 1)
    if s.f_rest.shape[1] == 45:
        sh = s.f_rest.astype(np.float64).reshape(n, 3, 15).transpose(0, 2, 1)  # [N, coefficient, channel]
        d = s.positions.astype(np.float64) - cam_model
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        x, y, z = d[:, 0:1], d[:, 1:2], d[:, 2:3]
        rgb = rgb + C1 * (-sh[:, 0] * y + sh[:, 1] * z - sh[:, 2] * x)
        rgb += (C2[0] * x * y * sh[:, 3] + C2[1] * y * z * sh[:, 4] + C2[2] * (2 * z * z - x * x - y * y) * sh[:, 5]
                + C2[3] * x * z * sh[:, 6] + C2[4] * (x * x - y * y) * sh[:, 7])
        rgb += (C3[0] * sh[:, 8] * (3 * x * x - y * y) * y + C3[1] * sh[:, 9] * x * y * z + C3[2] * sh[:, 10] * (4 * z * z - x * x - y * y) * y
                + C3[3] * sh[:, 11] * z * (2 * z * z - 3 * x * x - 3 * y * y) + C3[4] * sh[:, 12] * x * (4 * z * z - x * x - y * y)
                + C3[5] * sh[:, 13] * (x * x - y * y) * z + C3[6] * sh[:, 14] * x * (x * x - 3 * y * y))
    return rgb, alpha


def _front_end(s, fp):
    """View / clip transform, dist-stage cull, depth order (Appendix A.2). Returns (view [N,3], ndc [N,3], keep mask)."""
    V, P, M = _mat(fp.view), _mat(fp.proj), _mat(fp.model)
    p = np.concatenate([s.positions.astype(np.float64), np.ones((s.size(), 1))], 1)
    view = (V @ M @ p.T).T
    clip = (P @ view.T).T
    ndc = clip[:, :3] / clip[:, 3:4]
    lim = 1.0 + fp.frustum_dilation
    keep = ~((np.abs(ndc[:, 0]) > lim) | (np.abs(ndc[:, 1]) > lim) | (ndc[:, 2] < -fp.frustum_dilation) | (ndc[:, 2] > 1.0))
    return view[:, :3], ndc, keep


def _composite(order, frags, h, w, front_to_back):
    """Appendix A.7 on per-splat fragment lists: frags[i] = (ys, xs, alpha [K], rgb [3])."""
    img = np.zeros((h, w, 4))
    for i in order:
        if i not in frags:
            continue
        ys, xs, al, rgb = frags[i]
        if front_to_back:
            t = 1.0 - img[ys, xs, 3]
            img[ys, xs, :3] += rgb[None, :] * (al * t)[:, None]
            img[ys, xs, 3] += al * t
        else:
            img[ys, xs, :3] = rgb[None, :] * al[:, None] + img[ys, xs, :3] * (1 - al)[:, None]
            img[ys, xs, 3] += al
    return img


def _render_3dgs_float64(s, fp, front_to_back, ms_antialiasing=False):
    h, w = fp.height, fp.width
    view, ndc, keep = _front_end(s, fp)
    cam_model = (np.linalg.inv(_mat(fp.model)) @ np.array(list(fp.camera_position) + [1.0]))[:3]
    rgb, alpha = _colours(s, cam_model)
    R = _rotations(s.rotation)
    S = np.exp(s.scale.astype(np.float64))
    RS = R * S[:, None, :]
    cov3 = RS @ RS.transpose(0, 2, 1)
    Wm = (_mat(fp.view) @ _mat(fp.model))[:3, :3]
    fx, fy = fp.focal[0], fp.focal[1]
    frags = {}
    for i in np.nonzero(keep)[0]:
        if alpha[i] < fp.alpha_cull_threshold or not (0.0 <= ndc[i, 2] <= 1.0):
            continue
        x, y, z = view[i]
        J = np.array([[fx / z, 0, -fx * x / (z * z)], [0, fy / z, -fy * y / (z * z)]])
        T = J @ Wm
        c2 = T @ cov3[i] @ T.T
        a, b, d = c2[0, 0] + 0.3, c2[0, 1], c2[1, 1] + 0.3
        al_i = alpha[i]
        if ms_antialiasing:  # mip-splatting compensation: opacity *= sqrt(det / det') (threedgs.h.slang:63-76)
            al_i = al_i * np.sqrt(max((c2[0, 0] * c2[1, 1] - b * b) / (a * d - b * b), 0.0))
        m = 0.5 * (a + d)
        t = np.sqrt(max(0.1, m * m - (a * d - b * b)))
        l1, l2 = m + t, m - t
        if l2 <= 0:
            continue
        e1 = np.array([1.0 if abs(b) < 0.001 else b, l1 - a])
        e1 /= np.linalg.norm(e1)
        e2 = np.array([e1[1], -e1[0]])
        b1 = e1 * fp.splat_scale * min(np.sqrt(8.0) * np.sqrt(l1), 2048.0)
        b2 = e2 * fp.splat_scale * min(np.sqrt(8.0) * np.sqrt(l2), 2048.0)
        c = (ndc[i, :2] * 0.5 + 0.5) * np.array([w, h])
        ext = np.abs(b1) + np.abs(b2)
        x0, x1 = int(max(0, np.floor(c[0] - ext[0] - 1))), int(min(w - 1, np.ceil(c[0] + ext[0] + 1)))
        y0, y1 = int(max(0, np.floor(c[1] - ext[1] - 1))), int(min(h - 1, np.ceil(c[1] + ext[1] + 1)))
        if x1 < x0 or y1 < y0:
            continue
        yy, xx = np.mgrid[y0:y1 + 1, x0:x1 + 1]
        dx, dy = xx + 0.5 - c[0], yy + 0.5 - c[1]
        u = (dx * b1[0] + dy * b1[1]) / (b1 @ b1)
        v = (dx * b2[0] + dy * b2[1]) / (b2 @ b2)
        Aq = 8.0 * (u * u + v * v)
        al = np.exp(-0.5 * Aq) * al_i
        ok = (np.abs(u) <= 1) & (np.abs(v) <= 1) & (Aq <= 8.0) & (al > 1.0 / 255.0)
        if ok.any():
            frags[i] = (yy[ok], xx[ok], al[ok], rgb[i])
    z = ndc[:, 2]
    idx = np.nonzero(keep)[0]
    order = idx[np.argsort(z[idx] if front_to_back else -z[idx], kind="stable")]
    return _composite(order, frags, h, w, front_to_back)


def _agreement(oracle_img, ref_img, front_to_back):
    d = np.abs(oracle_img.astype(np.float64) - ref_img)
    if not front_to_back:
        d[..., 3] /= np.maximum(1.0, np.abs(ref_img[..., 3]))  # back-to-front alpha is a plain sum of alphas
    per_pixel = d.max(axis=-1)
    return per_pixel


def test_3dgs_oracle_frame_matches_float64_restatement_of_appendix_a():
    s = g.synth_scene(3000, 3, 0x3D65F064)
    s.scale += np.float32(1.0)  # larger footprints: more overlap per pixel
    cam = g.default_camera()
    w, h = 200, 150
    fp = O.frame_params(cam, w, h)
    pk = O.Packed(s)
    for ftb in (0, 1):
        img, keys, ids, _ = O.render(pk, fp, O.default_options(front_to_back=ftb))
        ref = _render_3dgs_float64(s, fp, bool(ftb))
        per_pixel = _agreement(img, ref, bool(ftb))
        covered = (ref[..., 3] > 0).mean()
        assert covered > 0.3, "the scene must cover a good part of the frame"
        # fp32 vs fp64: identical up to rounding everywhere, except where one fragment sits within rounding of a discard
        # threshold (alpha = 1/255, A = 8, quad edge) and is kept by one side only: such a flip changes a pixel by <= 1/255 * |c|
        assert np.quantile(per_pixel, 0.995) < 2e-5, np.quantile(per_pixel, 0.995)
        assert per_pixel.max() < 2.0 / 255.0, per_pixel.max()
        # the depth order the oracle sorted by == the float64 order, up to depths that collide in fp32
        _, ndc, keep = _front_end(s, fp)
        idx = np.nonzero(keep)[0]
        order = idx[np.argsort(ndc[idx, 2] if ftb else -ndc[idx, 2], kind="stable")]
        assert len(order) == len(ids) and (order == ids).mean() > 0.99
        z_sorted = ndc[ids, 2] if ftb else -ndc[ids, 2]
        assert np.all(np.diff(z_sorted) > -2e-7)


def _gut_response_float64(s, i, fp, px, py):
    """max over the ray of exp(-0.5 (x-mu)^T Sigma^-1 (x-mu)): the particle response stated with the covariance matrix
    instead of the reference's canonical-space cross product. Pinhole ray through (px+0.5, py+0.5) as the reference
    generates it (generatePinholeRay with its half-pixel sub-pixel offset on top of the pixel centre)."""
    w, h = fp.viewport[0], fp.viewport[1]
    vinv, pinv = _mat(fp.view_inverse), _mat(fp.proj_inverse)
    d = np.array([(px + 0.5) / w * 2 - 1, (py + 0.5) / h * 2 - 1, 1.0, 1.0])
    target = pinv @ d
    direction = (vinv @ np.array([target[0], target[1], target[2], 0.0]))[:3]
    direction /= np.linalg.norm(direction)
    origin = (vinv @ np.array([0, 0, 0, 1.0]))[:3]
    R = _rotations(s.rotation[i:i + 1])[0]
    S = np.exp(s.scale[i].astype(np.float64))
    cov = (R * S[None, :]) @ (R * S[None, :]).T
    prec = np.linalg.inv(cov)
    o = origin - s.positions[i].astype(np.float64)
    # minimise (o + t d)^T P (o + t d) over t
    t = -(direction @ prec @ o) / (direction @ prec @ direction)
    x = o + t * direction
    return np.exp(-0.5 * (x @ prec @ x))


def test_3dgut_fragment_matches_ray_gaussian_mahalanobis_distance():
    s = g.synth_scene(400, 0, 0x3D65F065)
    s.scale += np.float32(1.5)
    cam = g.default_camera()
    w, h = 160, 120
    fp = O.frame_params(cam, w, h)
    opt = O.default_gut_options(front_to_back=1)
    img, keys, ids, quads = O.render_gut(O.Packed(s), s.rotation, fp, opt, want_quads=True)
    alpha = 1.0 / (1.0 + np.exp(-s.opacity.astype(np.float64)))
    rng = np.random.default_rng(3)
    checked = accepted = 0
    for i in np.nonzero(quads["valid"] == 1)[0][:150]:
        cx, cy = quads[i]["center"]
        for _ in range(6):
            px = np.floor(cx + rng.uniform(-1, 1) * quads[i]["extent"][0]) + 0.5
            py = np.floor(cy + rng.uniform(-1, 1) * quads[i]["extent"][1]) + 0.5
            ok, op = O.gut_fragment(quads[i], px, py, fp, opt)
            resp = _gut_response_float64(s, i, fp, px, py)
            a = min(fp.alpha_clamp, resp * alpha[i])
            expect = a > 1 / 255 and resp > fp.kernel_min_response
            near = abs(a - 1 / 255) < 2e-4 or abs(resp - fp.kernel_min_response) < 2e-4
            checked += 1
            if near:
                continue
            assert ok == expect, (i, px, py, ok, op, a, resp)
            if ok:
                accepted += 1
                # the response is ill-conditioned in |camera - centre| / scale (a cancelling cross product in fp32):
                # allow the same 4e-8 |ro| the GPU parity test states
                ro = np.linalg.norm(s.positions[i].astype(np.float64) - np.array(cam.eye)) / np.exp(s.scale[i].min())
                assert abs(op - a) < 2e-5 + 4e-8 * ro, (op, a, ro)
    assert checked > 500 and accepted > 150


def test_3dgut_unscented_projection_centre_is_close_to_the_pinhole_projection():
    """The UT mean of the seven projected sigma points is a second-order estimate of the projected centre: it must
    coincide, up to a few per cent of the footprint, with the plain pinhole projection of the centre (an independent statement of where the
    quad has to be), and the UT covariance with the Jacobian-projected covariance of the 3DGS path."""
    s = g.synth_scene(500, 0, 0x3D65F066)
    cam = g.default_camera()
    w, h = 320, 240
    fp = O.frame_params(cam, w, h)
    img, keys, ids, quads = O.render_gut(O.Packed(s), s.rotation, fp, O.default_gut_options(), want_quads=True)
    view, ndc, keep = _front_end(s, fp)
    valid = np.nonzero(quads["valid"] == 1)[0]
    assert len(valid) > 400
    c = (ndc[valid, :2] * 0.5 + 0.5) * np.array([w, h])
    # (the UT mean carries the second-order term of the perspective map: a few per cent of the footprint, never more)
    assert (np.abs(quads["center"][valid] - c) / quads["extent"][valid]).max() < 0.05
    assert np.median(np.abs(quads["center"][valid] - c) / quads["extent"][valid]) < 0.01
    # extents: 3.33 sigma of the dilated Jacobian-projected covariance, capped by the opacity-limited bound
    R = _rotations(s.rotation)
    S = np.exp(s.scale.astype(np.float64))
    Wm = (_mat(fp.view) @ _mat(fp.model))[:3, :3]
    alpha = 1.0 / (1.0 + np.exp(-s.opacity.astype(np.float64)))
    errs = []
    for i in valid[:200]:
        x, y, z = view[i]
        J = np.array([[fp.focal[0] / z, 0, -fp.focal[0] * x / (z * z)], [0, fp.focal[1] / z, -fp.focal[1] * y / (z * z)]])
        T = J @ Wm
        RS = R[i] * S[i][None, :]
        c2 = T @ (RS @ RS.T) @ T.T
        ef = min(3.33, np.sqrt(2 * np.log(alpha[i] / 0.01)))
        ex, ey = ef * np.sqrt(c2[0, 0] + 0.3), ef * np.sqrt(c2[1, 1] + 0.3)
        errs.append(max(abs(quads[i]["extent"][0] - ex) / ex, abs(quads[i]["extent"][1] - ey) / ey))
    # (UT covariance vs first-order Jacobian covariance: second-order differences only)
    assert max(errs) < 0.08 and np.median(errs) < 0.01, (max(errs), np.median(errs))


def test_3dgut_fisheye_projection_and_ray_match_the_equidistant_model():
    """CAMERA_FISHEYE stated from the model instead of the shader: a point at angle theta from the optical axis lands
    focal * theta away from the principal point (equidistant projection), and the ray of a pixel at normalised radius r
    leaves at angle r * fov / 2. The oracle's fixed-sequence atan2 / acos / sin / cos must reproduce both."""
    s = g.synth_scene(500, 0, 0x3D65F067)
    cam = g.default_camera()
    cam.fov_deg = 120.0
    w, h = 320, 240
    fp = O.frame_params(cam, w, h, fisheye=True)
    opt = O.default_gut_options(camera_model=A.CAMERA_FISHEYE, front_to_back=1)
    img, keys, ids, quads = O.render_gut(O.Packed(s), s.rotation, fp, opt, want_quads=True)
    valid = np.nonzero(quads["valid"] == 1)[0]
    assert len(valid) > 300
    V, M = _mat(fp.view), _mat(fp.model)
    p = np.concatenate([s.positions.astype(np.float64), np.ones((s.size(), 1))], 1)
    view = (V @ M @ p.T).T[:, :3] * np.array([1.0, 1.0, -1.0])  # camera looks down +z in the sensor frame
    rho = np.hypot(view[:, 0], view[:, 1])
    theta = np.arctan2(rho, view[:, 2])
    c = np.stack([fp.focal[0] * view[:, 0] / rho * theta + w / 2, fp.focal[1] * view[:, 1] / rho * theta + h / 2], 1)
    rel = np.abs(quads["center"][valid] - c[valid]) / quads["extent"][valid]
    assert rel.max() < 0.08 and np.median(rel) < 0.01, (rel.max(), np.median(rel))
    # rays: opacity of a fragment == response of the float64 ray built from the model, for pixels across the frame
    vinv = _mat(fp.view_inverse)
    alpha = 1.0 / (1.0 + np.exp(-s.opacity.astype(np.float64)))
    R = _rotations(s.rotation)
    S = np.exp(s.scale.astype(np.float64))
    rng = np.random.default_rng(9)
    accepted = 0
    for i in valid[:150]:
        cx, cy = quads[i]["center"]
        for _ in range(4):
            px = np.floor(cx + rng.uniform(-0.6, 0.6) * quads[i]["extent"][0]) + 0.5
            py = np.floor(cy + rng.uniform(-0.6, 0.6) * quads[i]["extent"][1]) + 0.5
            u, v = px / (w - 1) * 2 - 1, py / (h - 1) * 2 - 1
            r = np.hypot(u, v)
            ok, op = O.gut_fragment(quads[i], px, py, fp, opt)
            if r > 1.0:
                assert not ok
                continue
            phi = np.arctan2(v, u)
            th = r * fp.fov_rad * 0.5
            d_cam = np.array([np.cos(phi) * np.sin(th), -np.sin(phi) * np.sin(th), -np.cos(th), 0.0])
            direction = (vinv @ d_cam)[:3]
            direction /= np.linalg.norm(direction)
            origin = (vinv @ np.array([0, 0, 0, 1.0]))[:3]
            RS = R[i] * S[i][None, :]
            prec = np.linalg.inv(RS @ RS.T)
            o = origin - s.positions[i].astype(np.float64)
            t = -(direction @ prec @ o) / (direction @ prec @ direction)
            x = o + t * direction
            resp = np.exp(-0.5 * (x @ prec @ x))
            a = min(fp.alpha_clamp, resp * alpha[i])
            if abs(a - 1 / 255) < 2e-4 or abs(resp - fp.kernel_min_response) < 2e-4:
                continue
            assert ok == (a > 1 / 255 and resp > fp.kernel_min_response), (i, px, py, ok, a, resp)
            if ok:
                accepted += 1
                ro = np.linalg.norm(o) / S[i].min()
                assert abs(op - a) < 5e-5 + 1e-7 * ro, (op, a, ro)
    assert accepted > 100


def test_3dgs_oracle_options_match_float64_restatement():
    """Mip-splatting antialiasing, a non-trivial model transform and a splat scale, through the same float64 restatement."""
    s = g.synth_scene(2000, 3, 0x3D65F068)
    s.scale += np.float32(0.8)
    cam = g.orbit_camera(1, 8)
    w, h = 180, 140
    fp = O.frame_params(cam, w, h)
    # model = translate * rotate(z, 0.4) * uniform scale 1.1, with its inverse (SplatSetDesc.transform / transformInverse)
    c, sn, k = np.cos(0.4), np.sin(0.4), 1.1
    Mm = np.array([[k * c, -k * sn, 0, 0.05], [k * sn, k * c, 0, -0.1], [0, 0, k, 0.02], [0, 0, 0, 1.0]])
    fp.model[:] = Mm.T.astype(np.float32).reshape(-1).tolist()
    fp.model_inverse[:] = np.linalg.inv(Mm).T.astype(np.float32).reshape(-1).tolist()
    fp.splat_scale = 0.75
    pk = O.Packed(s)
    for aa in (0, 1):
        img, keys, ids, _ = O.render(pk, fp, O.default_options(front_to_back=1, ms_antialiasing=aa))
        ref = _render_3dgs_float64(s, fp, True, ms_antialiasing=bool(aa))
        per_pixel = _agreement(img, ref, True)
        assert (ref[..., 3] > 0).mean() > 0.2
        assert np.quantile(per_pixel, 0.995) < 3e-5 and per_pixel.max() < 2.0 / 255.0, (np.quantile(per_pixel, 0.995), per_pixel.max())
