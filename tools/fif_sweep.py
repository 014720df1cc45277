import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import vk_gaussian_splatting_b200 as g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
s = g.synth_scene(n, 3, 0x3D650001)
fp = g.frame_params(g.default_camera(), w, h)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
r = g.GaussianSplatting(0, stream=stream.cuda_stream)
r.upload(s, g.default_options(front_to_back=1, transmittance_epsilon=2.0**-15))
for fif in (1, 2, 3, 4, 5, 6, 8):
    try:
        r.set_frames_in_flight(fif)
    except Exception:
        break
    r.set_frames_in_flight(fif)
    for _ in range(10): r.render_async(fp)
    r.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(K): r.render_async(fp)
    t_enq = (time.perf_counter() - t0) / K * 1e6
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"frames_in_flight={fif}: {ms*1000:.1f} us/frame  {1000/ms:.0f} fps   cpu enqueue {t_enq:.1f} us/frame", flush=True)
