"""A/B several tuning builds of the library in ONE GPU call.

    python tools/ab_variants.py build  name=DEF1,DEF2=3 name2=...     # in the build container: lib/libvkgs_b200_<name>.so
    python tools/ab_variants.py run [cfg2|cfg3|btf|gut ...]           # on the GPU box: one subprocess per library found

`run` prints one JSON line per (library, workload): frames/s with four frames in flight (host clock around 200
render_async calls + sync) and the per-kernel device times of a frame rendered alone."""
import json, os, subprocess, sys, time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def build(specs):
    from vk_gaussian_splatting_b200 import build as B
    for spec in specs:
        name, _, defs = spec.partition("=")
        B.build_variant(name, [d for d in defs.split(",") if d])
        print("built", name, defs)


def child(workloads):
    import numpy as np
    import vk_gaussian_splatting_b200 as g
    from vk_gaussian_splatting_b200 import _abi as A
    r = g.GaussianSplatting(0)
    W = {"cfg2": (1_000_000, 1920, 1080, 0x3D650001, dict(front_to_back=1, transmittance_epsilon=2.0 ** -15)),
         "cfg2x": (1_000_000, 1920, 1080, 0x3D650001, dict(front_to_back=1)),
         "btf": (1_000_000, 1920, 1080, 0x3D650001, dict(front_to_back=0)),
         "btfe": (1_000_000, 1920, 1080, 0x3D650001, dict(front_to_back=0, transmittance_epsilon=2.0 ** -15)),
         "cfg3": (6_000_000, 3840, 2160, 0x3D650002, dict(front_to_back=1, transmittance_epsilon=2.0 ** -15)),
         "cfg5": (30_000_000, 1920, 1080, 0x3D650004, dict(front_to_back=1, transmittance_epsilon=2.0 ** -15)),
         "gut": (1_000_000, 1920, 1080, 0x3D650001, dict(front_to_back=1, transmittance_epsilon=2.0 ** -15, pipeline=1)),
         "gutx": (1_000_000, 1920, 1080, 0x3D650001, dict(front_to_back=1, transmittance_epsilon=2.0 ** -15, pipeline=1, extent_projection=0)),
         "u8": (1_000_000, 1920, 1080, 0x3D650001, dict(front_to_back=1, transmittance_epsilon=2.0 ** -15, sh_format=2, rgba_format=2)),
         "f16": (1_000_000, 1920, 1080, 0x3D650001, dict(front_to_back=1, transmittance_epsilon=2.0 ** -15, sh_format=1, rgba_format=1))}
    scenes = {}
    for w in workloads:
        n, wd, ht, seed, kw = W[w]
        if (n, seed) not in scenes:
            scenes.clear()
            scenes[(n, seed)] = g.synth_scene(n, 3, seed)
        s = scenes[(n, seed)]
        r.upload(s, g.default_options(**kw))
        fp = g.frame_params(g.default_camera(), wd, ht)
        steps = 200 if n <= 1_000_000 else (60 if n <= 6_000_000 else 20)
        r.set_frames_in_flight(4)
        for _ in range(10):
            r.render_async(fp)
        r.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            r.render_async(fp)
        r.sync()
        dt = time.perf_counter() - t0
        r.set_frames_in_flight(1)
        t0 = time.perf_counter()
        for _ in range(50):
            r.render(fp, want_image=False)   # synchronous frames, one at a time, frame left on the device
        sync_us = (time.perf_counter() - t0) / 50 * 1e6
        r.set_profiling(True)
        acc = {}
        for _ in range(5):
            for _ in range(4):
                r.render_async(fp)
            st = r.last_frame_stats()
            for k, v in st.ms_kernel.items():
                acc[k] = acc.get(k, 0) + v / 5
        r.set_profiling(False)
        print(json.dumps({"lib": os.path.basename(os.environ.get("VKGS_LIB", "default")), "workload": w, "fps": round(steps / dt, 1), "sync_us": round(sync_us, 1),
                          "kernel_us": {k: round(v * 1000, 1) for k, v in acc.items() if v > 0.004}}), flush=True)
    r.close()


def run(workloads):
    libs = sorted((ROOT / "vk_gaussian_splatting_b200" / "lib").glob("libvkgs_b200*.so"))
    for lib in libs:
        env = dict(os.environ, VKGS_LIB=str(lib))
        subprocess.run([sys.executable, __file__, "child", *workloads], env=env)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "child":
        child(sys.argv[2:])
    else:
        run(sys.argv[2:] or ["cfg2"])
