"""One workload for ncu captures: python tools/prof_frame.py [fif=1|4] [workload: cfg2|u8|f16|cfg3|gut|btf] [frames]"""
import sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import vk_gaussian_splatting_b200 as g
fif = int(sys.argv[1]) if len(sys.argv) > 1 else 1
wl = sys.argv[2] if len(sys.argv) > 2 else "cfg2"
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kw = dict(front_to_back=1, transmittance_epsilon=2.0 ** -15)
n, w, h, seed = 1_000_000, 1920, 1080, 0x3D650001
if wl == "u8": kw.update(sh_format=2, rgba_format=2)
if wl == "f16": kw.update(sh_format=1, rgba_format=1)
if wl == "gut": kw.update(pipeline=1)
if wl == "btf": kw = dict(front_to_back=0)
if wl == "cfg3": n, w, h, seed = 6_000_000, 3840, 2160, 0x3D650002
s = g.synth_scene(n, 3, seed)
fp = g.frame_params(g.default_camera(), w, h)
r = g.GaussianSplatting(0)
r.upload(s, g.default_options(**kw))
r.set_frames_in_flight(fif)
for _ in range(frames):
    r.render_async(fp)
r.sync()
