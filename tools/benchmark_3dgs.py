#!/usr/bin/env python
"""The reference's benchmark_3dgs.cfg sequence on this path, logged in the reference's format.

    python tools/benchmark_3dgs.py [scene.ply|.spz|.splat] > _benchmark/log.txt
then the reference's benchmark.py parses the log (`parse_benchmark`). Without a scene file the
configs[1] synthetic scene is used. Sequences: SH storage fp32 / fp16 / uint8 (benchmark_3dgs.cfg
"--shformat 0/1/2"; the vert/mesh pipeline switch has no counterpart: both map onto the CUDA path).
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A, benchlog as B

scene = g.load_scene(sys.argv[1]) if len(sys.argv) > 1 else g.synth_scene(1_000_000, 3, 0x3D650001)
fp = g.frame_params(g.default_camera(), 1920, 1080)   # benchmark.py runs the app with --size 1920 1080
r = g.GaussianSplatting(0)
n = scene.size()
host_bytes = sum(a.nbytes for a in (scene.positions, scene.f_dc, scene.f_rest, scene.opacity, scene.scale, scene.rotation))
for i, (name, fmt) in enumerate((("CUDA pipeline fp32", A.FORMAT_FLOAT32), ("CUDA pipeline fp16", A.FORMAT_FLOAT16),
                                 ("CUDA pipeline uint8", A.FORMAT_UINT8))):
    r.upload(scene, g.default_options(sh_format=fmt))   # reference defaults otherwise (back-to-front)
    print(B.run_sequence(r, fp, i, name, f"--pipeline cuda\n--shformat {fmt}\n--updateData"), end="")
    el = (4, 2, 1)[fmt]
    scene_dev = n * (12 + 24 + 12 + 16 + (45 * el if scene.f_rest.shape[1] else 0))
    st = r.last_frame_stats()
    raster_dev = 2 * (4 * 4 * n + 48 * n + 16 * fp.width * fp.height) + 2 * 16 * max(8 * n, 1 << 20)
    print(B.memory_block(i, (host_bytes, scene_dev, scene_dev), (0, raster_dev, raster_dev)), end="")
r.close()
