"""Exploratory fuzz over options, frame parameters and cameras: CUDA path vs oracle (visible count, keys, ids bit-exact;
image within 1e-4 (+ the |ro| term on the 3DGUT pipeline)). python tools/fuzz_options.py [trials] [first trial]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
from oracle import oracle as O

def run(trials, first, r, log=print):
    bad = 0
    for t in range(first, first + trials):
        rng = np.random.default_rng(7000 + t)
        n = int(rng.choice([1, 2, 31, 257, 1000, 5000, 20000]))
        deg = int(rng.choice([0, 3]))
        s = g.synth_scene(n, deg, 0x3D65E000 + t)
        if rng.random() < 0.3:
            s.scale += np.float32(rng.uniform(-2.0, 2.5))   # tiny ... huge splats
        gut = rng.random() < 0.4
        kw = dict(front_to_back=int(rng.integers(0, 2)), ms_antialiasing=int(rng.integers(0, 2)),
                  frustum_culling_mode=int(rng.integers(0, 3)), disable_opacity_gaussian=int(rng.random() < 0.2))
        if gut:
            kw.update(pipeline=A.PIPELINE_3DGUT, extent_projection=int(rng.integers(0, 2)), kernel_degree=int(rng.choice([2, 2, 2, 0, 1, 3, 4, 5, 8])))
            if rng.random() < 0.25:
                kw["camera_model"] = A.CAMERA_FISHEYE
        else:
            kw.update(size_culling_mode=int(rng.random() < 0.3), point_cloud_mode=int(rng.random() < 0.15), show_sh_only=int(rng.random() < 0.15),
                      sh_format=int(rng.integers(0, 3)), rgba_format=int(rng.integers(0, 3)))
        cam = g.default_camera()
        mode = rng.integers(0, 4)
        if mode == 1:
            cam = g.orbit_camera(int(rng.integers(0, 8)), 8)
        elif mode == 2:   # inside the cloud, looking anywhere
            cam.eye[:] = tuple(rng.uniform(-0.8, 0.8, 3)); cam.ctr[:] = tuple(rng.uniform(-1, 1, 3))
        elif mode == 3:   # far away, narrow or wide
            cam.eye[:] = tuple(rng.uniform(-1, 1, 3) * rng.uniform(3, 40)); cam.ctr[:] = (0, 0, 0)
        cam.fov_deg = float(rng.choice([60.0, 60.0, 5.0, 20.0, 100.0, 150.0])) if not kw.get("camera_model") else float(rng.choice([60.0, 120.0, 170.0]))
        if rng.random() < 0.3:
            cam.znear, cam.zfar = float(rng.choice([0.001, 0.01, 1.0])), float(rng.choice([5.0, 100.0, 1e5]))
        w, h = [int(x) for x in rng.choice([(1, 1), (7, 5), (33, 32), (64, 48), (320, 200), (333, 217), (640, 97), (31, 400)])]
        fisheye = kw.get("camera_model") == A.CAMERA_FISHEYE
        tweak = {}
        if rng.random() < 0.5:
            tweak = dict(sh_degree=int(rng.integers(0, 4)), splat_scale=float(rng.choice([1.0, 0.3, 2.0])), frustum_dilation=float(rng.choice([0.2, 0.0, 0.5])),
                         size_culling_min_pixels=float(rng.choice([1.0, 3.0, 0.1])))
            if gut:
                tweak.update(alpha_cull_threshold=float(rng.choice([0.0, 1.0 / 255.0, 0.05])), kernel_min_response=float(rng.choice([0.0113, 0.0, 0.2])))
        desc = f"trial {t}: n={n} deg={deg} {w}x{h} fov={cam.fov_deg} mode={int(mode)} {kw} {tweak}"
        try:
            r.upload(s, g.default_options(**kw))
        except g.VkgsError as e:
            log(desc, "-> rejected by the product:", str(e)[:80], flush=True)
            continue
        fp, ofp = g.frame_params(cam, w, h, fisheye=fisheye), O.frame_params(cam, w, h, fisheye=fisheye)
        for k, v in tweak.items():
            setattr(fp, k, v); setattr(ofp, k, v)
        img, st, ids, keys = r.render(fp, want_sorted=True)
        okw = {k: v for k, v in kw.items() if k != "pipeline"}
        if gut:
            oimg, okeys, oids, _ = O.render_gut(O.Packed(s), s.rotation, ofp, O.default_gut_options(**okw))
        else:
            oimg, okeys, oids, _ = O.render(O.Packed(s, sh_format=kw["sh_format"], rgba_format=kw["rgba_format"]), ofp, O.default_options(**okw))
        msg = []
        if st.visible_count != len(oids):
            msg.append(f"visible {st.visible_count} vs {len(oids)}")
        elif not (np.array_equal(keys, okeys) and np.array_equal(ids, oids)):
            msg.append(f"keys equal {np.array_equal(keys, okeys)} ids equal {np.array_equal(ids, oids)}")
        if not np.array_equal(np.isfinite(img), np.isfinite(oimg)):
            msg.append("finite masks differ")
        d = np.abs(np.nan_to_num(img) - np.nan_to_num(oimg))
        if not kw["front_to_back"]:
            d[..., 3] /= np.maximum(1.0, np.abs(np.nan_to_num(oimg[..., 3])))
        tol = 1e-4
        if gut and len(oids):
            ro_max = float((np.linalg.norm(s.positions - np.array(cam.eye, np.float32), axis=1) / np.exp(s.scale.min(axis=1))).max())
            tol += 4e-8 * ro_max
        if d.max() > tol:
            msg.append(f"max diff {d.max():.3g} (tol {tol:.3g}) at {np.unravel_index(d.argmax(), d.shape)}")
        if msg:
            log(desc, "-> MISMATCH", "; ".join(msg), flush=True)
        bad += bool(msg)
    return bad


if __name__ == "__main__":
    n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    n_first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    n_bad = run(n_trials, n_first, g.GaussianSplatting(0))
    print(f"trials {n_first}..{n_first + n_trials - 1}: mismatching {n_bad}")
