"""All BASELINE.json GPU configurations + the sort-only sweep, one JSON line each (not the driver's
bench contract — that is bench.py; this feeds the tables in DESIGN.md / profiles/)."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))  # this checkout
import vk_gaussian_splatting_b200 as g

PEAK = 6538.6
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
r = g.GaussianSplatting(0, stream=stream.cuda_stream)

def run(name, n, w, h, seed, steps=100, fisheye=False, **optkw):
    t0 = time.time()
    s = g.synth_scene(n, 3, seed)
    t_gen = time.time() - t0
    t0 = time.time()
    r.upload(s, g.default_options(**{"front_to_back": 1, "transmittance_epsilon": 2.0 ** -15, **optkw}))
    t_up = time.time() - t0
    fp = g.frame_params(g.default_camera(), w, h, fisheye=fisheye)
    r.set_frames_in_flight(2)
    for _ in range(5): r.render_async(fp)
    r.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps): r.render_async(fp)
    e1.record(stream); torch.cuda.synchronize(); r.sync()
    ms = e0.elapsed_time(e1) / steps
    r.set_frames_in_flight(1); r.set_profiling(True)
    acc = {}
    for _ in range(5):
        for _ in range(4): r.render_async(fp)
        st = r.last_frame_stats()
        for k, v in st.ms_kernel.items(): acc[k] = acc.get(k, 0) + v / 5
    r.set_profiling(False)
    out = {"config": name, "splats": n, "size": [w, h], "fps": 1000 / ms, "ms_per_frame": ms, "msplats_per_s": n / ms / 1e3,
           "visible": st.visible_count, "tile_pairs": st.tile_pairs, "B_alg_MB": st.bytes_algorithmic / 1e6,
           "frame_hbm_gbs": st.bytes_algorithmic / ms / 1e6, "frame_hbm_frac": st.bytes_algorithmic / ms / 1e6 / PEAK,
           "kernel_us": {k: round(v * 1000, 1) for k, v in acc.items() if v > 0.004}, "gen_s": round(t_gen, 1), "upload_s": round(t_up, 1)}
    print(json.dumps(out), flush=True)

which = sys.argv[1:] or ["2", "3", "5", "sort"]
if "2" in which: run("cfg2 1M SH3 1080p", 1_000_000, 1920, 1080, 0x3D650001, 200)
if "3" in which: run("cfg3 6M SH3 4K", 6_000_000, 3840, 2160, 0x3D650002, 50)
if "5" in which: run("cfg5 30M SH3 1080p", 30_000_000, 1920, 1080, 0x3D650004, 20)
if "gut" in which: run("cfg2 1M SH3 1080p VK3DGUT pipeline", 1_000_000, 1920, 1080, 0x3D650001, 100, pipeline=1)
# the general 3DGUT blend instantiation (EXTENT_EIGEN quads / fisheye camera), same frame
if "gutx" in which:
    run("cfg2 1M SH3 1080p VK3DGUT, EXTENT_EIGEN", 1_000_000, 1920, 1080, 0x3D650001, 100, pipeline=1, extent_projection=0)
    run("cfg2 1M SH3 1080p VK3DGUT, fisheye camera", 1_000_000, 1920, 1080, 0x3D650001, 100, pipeline=1, camera_model=1, fisheye=True)
    run("cfg2 1M SH3 1080p back-to-front (reference default order)", 1_000_000, 1920, 1080, 0x3D650001, 100, front_to_back=0)
if "surf" in which: run("cfg2 1M SH3 1080p + surface info", 1_000_000, 1920, 1080, 0x3D650001, 100, surface_info=1)
if "sort" in which:
    rng = np.random.default_rng(5)
    for n in (1, 2, 4, 8, 16, 30):
        m = n * 1_000_000
        keys = rng.integers(0, 1 << 32, size=m, dtype=np.uint64).astype(np.uint32)
        vals = np.arange(m, dtype=np.uint32)
        k, v, ms = r.sort_pairs(keys, vals, repeats=5)
        ok = bool(np.all(k[1:] >= k[:-1]))
        print(json.dumps({"config": "sort-only", "pairs": m, "ms": ms, "gpairs_per_s": m / ms / 1e6, "GBps_at_68B": 68 * m / ms / 1e6,
                          "frac_of_peak": 68 * m / ms / 1e6 / PEAK, "sorted": ok}), flush=True)
