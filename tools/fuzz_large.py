"""Exploratory fuzz at sizes where the multi-partition machinery runs (look-back chains, huge-splat binning, tile-list regrowth):
100k-400k splats, up to 1280x720, splat sizes shifted from tiny to huge.  python tools/fuzz_large.py [trials] [first]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
from oracle import oracle as O


def run(trials, first, r, log=print):
    bad = 0
    for t in range(first, first + trials):
        rng = np.random.default_rng(13000 + t)
        n = int(rng.choice([100_000, 200_000, 400_000]))
        s = g.synth_scene(n, int(rng.choice([0, 3])), 0x3D65B000 + t)
        shift = float(rng.choice([0.0, -1.5, 1.0, 2.0, 2.5]))
        s.scale += np.float32(shift)
        if rng.random() < 0.5:  # a handful of screen-filling splats
            s.scale[rng.integers(0, n, 20)] = np.float32(np.log(rng.uniform(0.3, 3.0)))
        gut = rng.random() < 0.3 and shift <= 1.0
        kw = dict(front_to_back=int(rng.integers(0, 2)), ms_antialiasing=int(rng.integers(0, 2)), frustum_culling_mode=int(rng.integers(0, 3)))
        if gut:
            kw["pipeline"] = A.PIPELINE_3DGUT
        cam = g.orbit_camera(int(rng.integers(0, 8)), 8) if rng.random() < 0.5 else g.default_camera()
        if rng.random() < 0.3:
            cam.eye[:] = tuple(rng.uniform(-0.6, 0.6, 3))
        w, h = [int(x) for x in rng.choice([(1280, 720), (960, 540), (1000, 1000), (1920, 300)])]
        r.upload(s, g.default_options(**kw))
        t0 = time.time()
        img, st, ids, keys = r.render(g.frame_params(cam, w, h), want_sorted=True)
        okw = {k: v for k, v in kw.items() if k != "pipeline"}
        if gut:
            oimg, okeys, oids, _ = O.render_gut(O.Packed(s), s.rotation, O.frame_params(cam, w, h), O.default_gut_options(**okw))
        else:
            oimg, okeys, oids, _ = O.render(O.Packed(s), O.frame_params(cam, w, h), O.default_options(**okw))
        msg = []
        if st.visible_count != len(oids) or not (np.array_equal(keys, okeys) and np.array_equal(ids, oids)):
            msg.append(f"visible {st.visible_count} vs {len(oids)} / keys / ids differ")
        d = np.abs(img - oimg)
        if not kw["front_to_back"]:
            d[..., 3] /= np.maximum(1.0, np.abs(oimg[..., 3]))
        tol = 1e-4
        if gut:
            tol += 4e-8 * float((np.linalg.norm(s.positions - np.array(cam.eye, np.float32), axis=1) / np.exp(s.scale.min(axis=1))).max())
        if not (d.max() <= tol):
            msg.append(f"max diff {d.max():.3g} (tol {tol:.3g})")
        log(f"trial {t}: n={n} shift={shift} {w}x{h} {kw} pairs={st.tile_pairs} {time.time() - t0:.1f}s ->", "MISMATCH " + "; ".join(msg) if msg else "ok", flush=True)
        bad += bool(msg)
    return bad


if __name__ == "__main__":
    n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    n_first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print(f"mismatching {run(n_trials, n_first, g.GaussianSplatting(0))} of {n_trials}")
