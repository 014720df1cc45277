"""Exploratory stress of tile-list regrowth with frames in flight: bursts of asynchronous frames to pinned host memory whose
cameras alternate between few and very many tile pairs (every burst starts from a freshly uploaded scene, i.e. from the
initial tile-list capacity); every frame of every burst must equal the synchronous frame of a second context.
python tools/fuzz_overflow.py [bursts]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import vk_gaussian_splatting_b200 as g


def run(bursts, r, fresh, log=print):
    bad = 0
    rng = np.random.default_rng(777)
    for b in range(bursts):
        n = int(rng.choice([20_000, 60_000]))
        s = g.synth_scene(n, 0, 0x3D65A000 + b)
        s.scale += np.float32(rng.choice([0.0, 1.0, 2.0]))
        opt = g.default_options(front_to_back=int(rng.integers(0, 2)))
        w, h = [int(x) for x in rng.choice([(640, 360), (320, 200), (960, 540)])]
        cams = []
        for _ in range(int(rng.integers(2, 9))):
            cam = g.orbit_camera(int(rng.integers(0, 8)), 8)
            if rng.random() < 0.5:  # inside the cloud: huge splats, many more pairs
                cam.eye[:] = tuple(rng.uniform(-0.3, 0.3, 3))
            cams.append(cam)
        fif = int(rng.integers(1, 5))
        r.upload(s, opt)
        r.set_frames_in_flight(fif)
        # half of the bursts change the viewport from frame to frame (the slot's targets are reallocated under frames in flight)
        sizes = [(w, h)] * len(cams) if rng.random() < 0.5 else [tuple(int(x) for x in rng.choice([(640, 360), (320, 200), (960, 540), (333, 217)])) for _ in cams]
        bufs = [torch.zeros((hh, ww, 4), dtype=torch.float32, pin_memory=True).numpy() for (ww, hh) in sizes]
        try:
            for cam, buf, (ww, hh) in zip(cams, bufs, sizes):
                r.render_to_host_async(g.frame_params(cam, ww, hh), buf)
            # every call that ends a batch must leave the host destinations complete: vkgs_sync, or one of the calls that
            # change state under frames in flight (they complete the pending frames first, in the old state)
            how = int(rng.integers(0, 4))
            if how == 0:
                r.sync()
            elif how == 1:
                r.set_frames_in_flight(int(rng.integers(1, 5)))
            elif how == 2:
                r.set_target_format(1)   # RGBA16F from now on: the pending fp32 frames must still arrive as fp32
                r.set_target_format(0)
            else:
                r.upload(g.synth_scene(1000, 0, 1), opt)
            lost = False
        except g.VkgsError as e:
            # documented: a slot reused before vkgs_sync by a frame after an overflowing one cannot be repaired
            lost = True
            log(f"burst {b}: sync reported {str(e)[:70]} (fif {fif}, {len(cams)} frames)")
            for cam, buf, (ww, hh) in zip(cams, bufs, sizes):
                r.render_to_host_async(g.frame_params(cam, ww, hh), buf)
            r.sync()
        fresh.upload(s, opt)
        for i, (cam, buf, (ww, hh)) in enumerate(zip(cams, bufs, sizes)):
            want = fresh.render(g.frame_params(cam, ww, hh))[0]
            if not np.array_equal(buf, want):
                bad += 1
                log(f"burst {b}: frame {i} of {len(cams)} differs (fif {fif}, n {n}, {w}x{h}, lost={lost}, max diff {np.abs(buf - want).max():.3g})", flush=True)
    r.set_frames_in_flight(1)
    return bad


if __name__ == "__main__":
    n_b = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    print(f"differing frames: {run(n_b, g.GaussianSplatting(0), g.GaussianSplatting(0))}")
