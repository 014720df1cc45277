import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import vk_gaussian_splatting_b200 as g
s = g.synth_scene(1_000_000, 3, 0x3D650001)
r = g.GaussianSplatting(0)
def run(name, eps, shdeg, w=1920, h=1080):
    r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=eps))
    fp = g.frame_params(g.default_camera(), w, h); fp.sh_degree = shdeg
    r.set_frames_in_flight(4)
    for _ in range(10): r.render_async(fp)
    r.sync()
    t0 = time.perf_counter()
    for _ in range(300): r.render_async(fp)
    r.sync()
    dt = (time.perf_counter() - t0) / 300
    r.set_frames_in_flight(1); r.set_profiling(True)
    for _ in range(4): r.render_async(fp)
    st = r.last_frame_stats(); r.set_profiling(False)
    k = {a: round(b * 1000, 1) for a, b in st.ms_kernel.items() if b > 0.004}
    print(f"{name}: {dt*1e6:.1f} us/frame pipelined; alone sum {sum(k.values()):.0f} us {k}", flush=True)
run("normal", 2.0**-15, 3)
run("sh0", 2.0**-15, 0)
run("eps0.5 (blend cut short)", 0.5, 3)
run("eps0.5 + sh0", 0.5, 0)
run("tiny viewport 480x270 (raster shrinks, front end same)", 2.0**-15, 3, 480, 270)
