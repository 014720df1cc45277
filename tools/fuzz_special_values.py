"""Exploratory fuzz: small scenes with special values injected (NaN / inf / zero / denormal / huge) in every attribute,
CUDA path vs oracle: visible count, keys, ids bit-exact; image equal within 1e-4 where both are finite, non-finite in the
same places. python tools/fuzz_special_values.py [trials] [pipeline 0|1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
from oracle import oracle as O

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 20
pipeline = int(sys.argv[2]) if len(sys.argv) > 2 else 0
attrs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["positions", "f_dc", "f_rest", "opacity", "scale", "rotation"]

FINITE = np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 3e38, -3e38, 1e20, -1e20, 1e-20, 88.0, -88.0, 104.0, -104.0, 17.0, -17.0, 5.5, -5.5], np.float32)
SPECIAL = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1e-45, -1e-45, 1e-38, 3e38, -3e38, 1e20, -1e20, 1e-20, 88.0, -88.0, 104.0, -104.0], np.float32)
if len(sys.argv) > 4 and sys.argv[4] == "finite":
    SPECIAL = FINITE
r = g.GaussianSplatting(0)
bad = 0
for t in range(trials):
    rng = np.random.default_rng(1000 + t)
    s = g.synth_scene(3000, 3, 0x3D65F000 + t)
    for name in attrs:
        arr = getattr(s, name)
        flat = arr.reshape(-1)
        k = int(rng.integers(0, 40))
        idx = rng.integers(0, flat.size, k)
        flat[idx] = SPECIAL[rng.integers(0, SPECIAL.size, k)]
    ftb = int(t & 1)
    kw = dict(front_to_back=ftb)
    if pipeline:
        kw["pipeline"] = A.PIPELINE_3DGUT
    if t % 3 == 0:
        kw["ms_antialiasing"] = 1
    if t % 5 == 0 and not pipeline:
        kw["size_culling_mode"] = 1
    cam = g.orbit_camera(t % 8, 8)
    w, h = 320, 200
    r.upload(s, g.default_options(**kw))
    fp = g.frame_params(cam, w, h)
    img, st, ids, keys = r.render(fp, want_sorted=True)
    if pipeline:
        oimg, okeys, oids, _ = O.render_gut(O.Packed(s), s.rotation, O.frame_params(cam, w, h), O.default_gut_options(**{k: v for k, v in kw.items() if k != "pipeline"}))
    else:
        oimg, okeys, oids, _ = O.render(O.Packed(s), O.frame_params(cam, w, h), O.default_options(**kw))
    msg = []
    if st.visible_count != len(oids):
        msg.append(f"visible {st.visible_count} vs {len(oids)}")
    elif not (np.array_equal(keys, okeys) and np.array_equal(ids, oids)):
        msg.append(f"keys equal {np.array_equal(keys, okeys)} ids equal {np.array_equal(ids, oids)}")
        j = int(np.argmax((keys != okeys) | (ids != oids)))
        msg.append(f"first at rank {j}: gpu ({ids[j]}, {keys[j]:#x}) oracle ({oids[j]}, {okeys[j]:#x}); pos[gpu id] {s.positions[ids[j]]} pos[oracle id] {s.positions[oids[j]]}")
    fin_g, fin_o = np.isfinite(img), np.isfinite(oimg)
    if not np.array_equal(fin_g, fin_o):
        msg.append(f"finite masks differ at {int((fin_g != fin_o).sum())} values (gpu non-finite {int((~fin_g).sum())}, oracle {int((~fin_o).sum())})")
    both = fin_g & fin_o
    d = np.abs(np.where(both, img, 0) - np.where(both, oimg, 0))
    if not ftb:
        d[..., 3] /= np.maximum(1.0, np.abs(np.where(both[..., 3], oimg[..., 3], 0)))
    d = d / np.maximum(1.0, np.abs(np.where(both, oimg, 0)))  # huge colours: relative
    if d.max() > 1e-4:
        msg.append(f"max diff {d.max():.3g} at {np.unravel_index(d.argmax(), d.shape)}")
    print(f"trial {t} {kw}: " + ("ok" if not msg else "MISMATCH " + "; ".join(msg)), flush=True)
    bad += bool(msg)
print("mismatching trials:", bad)
