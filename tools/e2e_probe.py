"""End-to-end frames/s of the asynchronous path (frame copied to pinned host memory every step, four frames in flight):
python tools/e2e_probe.py [f16|f32|u8] [frames]   (VKGS_LIB selects a tuning build)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
fmt = sys.argv[1] if len(sys.argv) > 1 else "f16"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 600
F = {"f16": (A.FORMAT_FLOAT16, torch.float16), "f32": (A.FORMAT_FLOAT32, torch.float32), "u8": (A.FORMAT_UINT8, torch.uint8)}[fmt]
s = g.synth_scene(1_000_000, 3, 0x3D650001)
r = g.GaussianSplatting(0)
r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15, target_format=F[0]))
w, h = 1920, 1080
fp = g.frame_params(g.default_camera(), w, h)
host = [torch.empty((h, w, 4), dtype=F[1], pin_memory=True).numpy() for _ in range(4)]
r.set_frames_in_flight(4)
for rep in range(3):
    for i in range(20):
        r.render_to_host_async(fp, host[i % 4])
    r.sync()
    t0 = time.perf_counter()
    for i in range(frames):
        r.render_to_host_async(fp, host[i % 4])
    r.sync()
    dt = time.perf_counter() - t0
    print(f"{os.path.basename(os.environ.get('VKGS_LIB', 'default'))} {fmt}: {frames / dt:.1f} frames/s, {host[0].nbytes * frames / dt / 1e9:.1f} GB/s", flush=True)
