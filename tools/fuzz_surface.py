"""Exploratory fuzz of the NEED_SURFACE_INFO side outputs (normals, picked depth + transmittance, splat id) against the
oracle: random scenes with flat / needle particles, cameras, viewports, options.  python tools/fuzz_surface.py [trials] [first]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vk_gaussian_splatting_b200 as g
from oracle import oracle as O


def run(trials, first, r, log=print):
    bad = 0
    for t in range(first, first + trials):
        rng = np.random.default_rng(11000 + t)
        n = int(rng.choice([1, 50, 1000, 8000, 20000]))
        s = g.synth_scene(n, int(rng.choice([0, 3])), 0x3D65C000 + t)
        if rng.random() < 0.7:
            s.scale[::int(rng.integers(3, 50)), int(rng.integers(0, 3))] = np.log(np.float32(10.0 ** rng.uniform(-9, -3)))
        if rng.random() < 0.5:
            s.scale[::int(rng.integers(5, 90)), :2] = np.log(np.float32(10.0 ** rng.uniform(-9, -3)))
        if rng.random() < 0.3:
            s.scale += np.float32(rng.uniform(-1.5, 2.0))
        if os.environ.get("FUZZ_SPECIAL"):
            SPECIAL = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-45, 1e-38, 3e38, -3e38, 1e20, -1e20, 88.0, -88.0, 104.0, -104.0], np.float32)
            for name in ("positions", "scale", "rotation"):
                flat = getattr(s, name).reshape(-1)
                k = int(rng.integers(1, 30))
                flat[rng.integers(0, flat.size, k)] = SPECIAL[rng.integers(0, SPECIAL.size, k)]
        kw = dict(front_to_back=1, quantize_normals=int(rng.integers(0, 2)), ms_antialiasing=int(rng.integers(0, 2)),
                  frustum_culling_mode=int(rng.integers(0, 3)))
        cam = g.orbit_camera(int(rng.integers(0, 8)), 8) if rng.random() < 0.5 else g.default_camera()
        if rng.random() < 0.3:
            cam.eye[:] = tuple(rng.uniform(-0.8, 0.8, 3))
        w, h = [int(x) for x in rng.choice([(320, 200), (333, 217), (64, 48), (640, 97), (7, 5)])]
        tweak = {}
        if rng.random() < 0.5:
            tweak = dict(depth_iso_threshold=float(rng.choice([0.7, 0.5, 0.95, 0.1])), thin_particle_threshold=float(rng.choice([0.005, 0.0005, 0.05])))
        fp, ofp = g.frame_params(cam, w, h), O.frame_params(cam, w, h)
        for k, v in tweak.items():
            setattr(fp, k, v); setattr(ofp, k, v)
        iso = float(fp.depth_iso_threshold)
        r.upload(s, g.default_options(surface_info=1, **kw))
        img, st, ids, _ = r.render(fp, want_sorted=True)
        nrm, dt, sid = r.read_surface_info(w, h)
        oimg, onrm, odt, osid, oids = O.render_surface(O.Packed(s), s.rotation, ofp, O.default_options(**kw))
        msg = []
        if not np.array_equal(ids, oids):
            msg.append("ids differ")
        if not (np.abs(img - oimg).max() <= 1e-4):
            msg.append(f"colour diff {np.abs(img - oimg).max():.3g}")
        near_iso = np.abs(odt[..., 1] - iso) < 1e-5
        # normals are integrated up to the pick: a pixel whose pick flips (T within 1e-6 of the threshold) may differ
        nd = np.abs(nrm - onrm).max(axis=-1)
        tol_n = 1e-4 if not kw["quantize_normals"] else 1e-4
        if not (nd[~near_iso] <= tol_n).all():
            msg.append(f"normal diff {nd[~near_iso].max():.3g} at {np.unravel_index(np.where(near_iso, 0, nd).argmax(), nd.shape)}")
        if not (np.abs(dt[..., 1] - odt[..., 1]).max() <= 1e-4):
            msg.append(f"transmittance diff {np.abs(dt[..., 1] - odt[..., 1]).max():.3g}")
        same_pick = (dt[..., 0] == odt[..., 0]) | near_iso
        if not same_pick.all():
            msg.append(f"depth pick differs at {int((~same_pick).sum())} pixels")
        same_id = (sid == osid) | near_iso
        if not same_id.all():
            msg.append(f"splat id differs at {int((~same_id).sum())} pixels")
        if msg:
            log(f"trial {t}: n={n} {w}x{h} {kw} {tweak} -> MISMATCH", "; ".join(msg), flush=True)
        bad += bool(msg)
    return bad


if __name__ == "__main__":
    n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    n_first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print(f"mismatching {run(n_trials, n_first, g.GaussianSplatting(0))} of {n_trials}")
