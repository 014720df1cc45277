"""Frames/s of short and long bursts of asynchronous frames (pipeline fill / drain included, as in `bench.py --steps K`):
python tools/fill_probe.py   (VKGS_NO_IDLE_FULL / VKGS_THIN_MIN_INFLIGHT select the policy under test)"""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_gaussian_splatting_b200 as g
s = g.synth_scene(1_000_000, 3, 0x3D650001)
r = g.GaussianSplatting(0)
r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15))
fp = g.frame_params(g.default_camera(), 1920, 1080)
r.set_frames_in_flight(4)
for _ in range(50):
    r.render_async(fp)
r.sync()
out = {}
for K, reps in ((20, 40), (100, 10), (1000, 3)):
    v = []
    for _ in range(reps):
        r.sync()
        t0 = time.perf_counter()
        for _ in range(K):
            r.render_async(fp)
        r.sync()
        v.append(K / (time.perf_counter() - t0))
    out[K] = (statistics.median(v), min(v), max(v))
t0 = time.perf_counter()
r.set_frames_in_flight(1)
print(" ".join(f"K={k}: {m:.0f} fps [{lo:.0f}..{hi:.0f}]" for k, (m, lo, hi) in out.items()),
      f"| policy NO_IDLE_FULL={os.environ.get('VKGS_NO_IDLE_FULL', '0')} THIN_MIN={os.environ.get('VKGS_THIN_MIN_INFLIGHT', 'default')}", flush=True)
