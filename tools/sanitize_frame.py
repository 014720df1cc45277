"""A few small frames through every kernel family, for compute-sanitizer (tools/sanitize.sh):
3DGS front-to-back / back-to-front, uint8 + fp16 storage, surface info, fragment counters, multi-instance,
3DGUT (pinhole CONIC, EIGEN, fisheye), caller-ordered frames, the stand-alone sorts, four frames in flight, metrics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A

n, w, h = int(os.environ.get("SAN_SPLATS", "20000")), 320, 200
s = g.synth_scene(n, 3, 0x3D65AA01)
cam = g.default_camera()
fp = g.frame_params(cam, w, h)
r = g.GaussianSplatting(0)
frames = 0
for kw in (dict(front_to_back=1, transmittance_epsilon=2.0 ** -15), dict(front_to_back=0), dict(front_to_back=1, sh_format=2, rgba_format=2),
           dict(front_to_back=1, sh_format=1, rgba_format=1, target_format=A.FORMAT_FLOAT16), dict(front_to_back=1, surface_info=1),
           dict(front_to_back=1, size_culling_mode=1), dict(front_to_back=0, ms_antialiasing=1, target_format=A.FORMAT_UINT8),
           dict(front_to_back=1, pipeline=A.PIPELINE_3DGUT), dict(front_to_back=0, pipeline=A.PIPELINE_3DGUT, extent_projection=A.EXTENT_EIGEN),
           dict(front_to_back=1, pipeline=A.PIPELINE_3DGUT, camera_model=A.CAMERA_FISHEYE)):
    r.upload(s, g.default_options(**kw))
    f = g.frame_params(cam, w, h, fisheye=kw.get("camera_model") == A.CAMERA_FISHEYE)
    img, st, ids, keys = r.render(f, want_sorted=True)
    assert st.visible_count > 0 and np.isfinite(np.asarray(img, np.float32)).all()
    r.set_frames_in_flight(4)
    for _ in range(4):
        r.render_async(f)
    r.sync()
    r.set_frames_in_flight(1)
    frames += 5
# caller-ordered frame, stand-alone sorts
r.upload(s, g.default_options(front_to_back=0))
order = np.random.default_rng(1).permutation(n).astype(np.uint32)
r.render_presorted(fp, order)
k = np.random.default_rng(2).integers(0, 2 ** 32, 100_000, dtype=np.uint64).astype(np.uint32)
v = np.arange(k.size, dtype=np.uint32)
ks, vs, _ = r.sort_pairs(k, v)
assert np.all(np.diff(ks.astype(np.int64)) >= 0)
r.close()
print(f"sanitize_frame ok: {frames + 1} frames, {n} splats, {w}x{h}")
