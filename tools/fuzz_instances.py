"""Exploratory fuzz of multi-instance scenes: 1-8 instances of 1-3 splat sets with random rigid + uniform / non-uniform scale
transforms, random options, CUDA path vs oracle (3DGS pipeline: O.render_scene).  python tools/fuzz_instances.py [trials] [first]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
from oracle import oracle as O


def random_transform(rng):
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    sc = np.full(3, rng.uniform(0.3, 1.5)) if rng.random() < 0.7 else rng.uniform(0.4, 1.4, 3)
    if rng.random() < 0.1:
        sc[0] = -sc[0]  # mirrored
    m = np.eye(4)
    m[:3, :3] = R @ np.diag(sc)
    m[:3, 3] = rng.uniform(-1.2, 1.2, 3)
    return np.ascontiguousarray(m.T, np.float32), np.ascontiguousarray(np.linalg.inv(m).T, np.float32)


def run(trials, first, r, log=print):
    bad = 0
    for t in range(first, first + trials):
        rng = np.random.default_rng(9000 + t)
        nsets = int(rng.integers(1, 4))
        sets = [g.synth_scene(int(rng.choice([1, 77, 1000, 4099, 9000])), int(rng.choice([0, 3])), 0x3D65D000 + 16 * t + k) for k in range(nsets)]
        inst = [(int(rng.integers(0, nsets)), *random_transform(rng)) for _ in range(int(rng.integers(1, 9)))]
        kw = dict(front_to_back=int(rng.integers(0, 2)), ms_antialiasing=int(rng.integers(0, 2)), frustum_culling_mode=int(rng.integers(0, 3)))
        gut = rng.random() < 0.35 and len(inst) <= 8
        if gut:
            kw.update(pipeline=A.PIPELINE_3DGUT, extent_projection=int(rng.integers(0, 2)))
        cam = g.orbit_camera(int(rng.integers(0, 8)), 8) if rng.random() < 0.5 else g.default_camera()
        w, h = [int(x) for x in rng.choice([(320, 200), (333, 217), (64, 48), (640, 97)])]
        r.upload_scene(sets, inst, g.default_options(**kw))
        img, st, ids, keys = r.render(g.frame_params(cam, w, h), want_sorted=True)
        if gut:
            okw = {k: v for k, v in kw.items() if k != "pipeline"}
            oimg, okeys, oids = O.render_scene([O.Packed(s) for s in sets], inst, O.frame_params(cam, w, h), O.default_gut_options(**okw),
                                               rotations=[s.rotation for s in sets])
        else:
            oimg, okeys, oids = O.render_scene([O.Packed(s) for s in sets], inst, O.frame_params(cam, w, h), O.default_options(**kw))
        msg = []
        if st.visible_count != len(oids):
            msg.append(f"visible {st.visible_count} vs {len(oids)}")
        elif not (np.array_equal(keys, okeys) and np.array_equal(ids, oids)):
            msg.append("keys / ids differ")
        d = np.abs(img - oimg)
        if not kw["front_to_back"]:
            d[..., 3] /= np.maximum(1.0, np.abs(oimg[..., 3]))
        tol = 1e-4
        if gut:  # |ro| = distance / scale in the instance's model space (the instance scale cancels)
            tol += 4e-8 * max(float(20.0 / np.exp(s.scale.min())) for s in sets)
        if not (d.max() <= tol):
            msg.append(f"max diff {d.max():.3g}")
        if msg:
            log(f"trial {t}: sets {[s.size() for s in sets]} instances {[i[0] for i in inst]} {w}x{h} {kw} -> MISMATCH", "; ".join(msg), flush=True)
        bad += bool(msg)
    return bad


if __name__ == "__main__":
    n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    n_first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print(f"mismatching {run(n_trials, n_first, g.GaussianSplatting(0))} of {n_trials}")
