"""Robustness probe: camera inside the 1M-splat scene (huge near splats, long tile lists, list regrow)."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import vk_gaussian_splatting_b200 as g
s = g.synth_scene(1_000_000, 3, 0x3D650001)
r = g.GaussianSplatting(0)
for name, opt in (("ftb eps", g.default_options(front_to_back=1, transmittance_epsilon=2.0 ** -15)), ("btf exact", g.default_options())):
    r.upload(s, opt)
    for eye in ((0.05, 0.02, 0.03), (0.6, 0.1, -0.4)):
        fp = g.frame_params(g.make_camera(eye, ctr=(1, 0.2, 0.3)), 1920, 1080)
        t0 = time.time(); img, st, _, _ = r.render(fp); t1 = time.time() - t0
        r.set_frames_in_flight(1); r.set_profiling(True)
        for _ in range(3): r.render_async(fp)
        st = r.last_frame_stats(); r.set_profiling(False); r.set_frames_in_flight(4)
        print(name, eye, "first render s", round(t1, 3), "V", st.visible_count, "pairs", st.tile_pairs, "ms_total", round(st.ms_total, 3),
              {k: round(v * 1000) for k, v in st.ms_kernel.items() if v > 0.004}, "finite", bool(np.isfinite(img).all()), flush=True)
