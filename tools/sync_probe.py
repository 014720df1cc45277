"""Latency of the synchronous call into pinned host memory: python tools/sync_probe.py  (VKGS_NO_STRIPS=1 for the one-piece copy)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A
s = g.synth_scene(1_000_000, 3, 0x3D650001)
r = g.GaussianSplatting(0)
for name, fmt, dt in (("f16", A.FORMAT_FLOAT16, torch.float16), ("f32", A.FORMAT_FLOAT32, torch.float32), ("u8", A.FORMAT_UINT8, torch.uint8)):
    for w, h in ((1920, 1080), (3840, 2160)):
        r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15, target_format=fmt))
        fp = g.frame_params(g.default_camera(), w, h)
        host = torch.empty((h, w, 4), dtype=dt, pin_memory=True).numpy()
        for _ in range(10):
            r.render(fp, out=host)
        t0 = time.perf_counter()
        for _ in range(100):
            r.render(fp, out=host)
        dt_us = (time.perf_counter() - t0) / 100 * 1e6
        ref = host.copy()
        os.environ["X"] = "1"
        print(f"{name} {w}x{h}: {dt_us:.1f} us per synchronous call (strips {'off' if os.environ.get('VKGS_NO_STRIPS') == '1' else 'on'}), checksum {int(ref.view(np.uint8).astype(np.uint64).sum())}", flush=True)
