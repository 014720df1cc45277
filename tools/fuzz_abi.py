"""Garbage in the option and frame-parameter structures (invalid enums, NaN / inf / huge matrices and scalars, zero and absurd
sizes): every call must return an error code or a frame — no crash, no hang, and the context must keep working afterwards.
python tools/fuzz_abi.py [trials]   (run under `timeout`)"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vk_gaussian_splatting_b200 as g
from vk_gaussian_splatting_b200 import _abi as A


def run(trials, r, log=print):
    rng = np.random.default_rng(31337)
    s = g.synth_scene(5000, 3, 0x3D659001)
    cam = g.default_camera()
    good_fp = g.frame_params(cam, 160, 100)
    r.upload(s, g.default_options(front_to_back=1))
    want = r.render(good_fp)[0].copy()
    errors = frames = 0
    WEIRD_F = [np.nan, np.inf, -np.inf, 0.0, -1.0, 1e30, -1e30, 1e-30, 3e38]
    WEIRD_U = [0, 1, 2, 3, 4, 7, 255, 65535, 65536, 2**31 - 1, 2**32 - 1]
    for t in range(trials):
        opt = g.default_options(front_to_back=int(rng.integers(0, 2)))
        if rng.random() < 0.6:
            for name, ctype in A.Options._fields_:
                if name.startswith("_"):
                    continue
                if rng.random() < 0.25:
                    setattr(opt, name, float(rng.choice(WEIRD_F)) if ctype is C.c_float else int(rng.choice(WEIRD_U)))
        try:
            r.upload(s, opt)
            uploaded = True
        except g.VkgsError:
            errors += 1
            uploaded = False
            r.upload(s, g.default_options(front_to_back=1))
        fp = g.frame_params(cam, int(rng.choice([160, 1, 33])), int(rng.choice([100, 1, 17])))
        for name, ctype in A.FrameParams._fields_:
            if rng.random() < 0.15:
                if ctype is C.c_float:
                    setattr(fp, name, float(rng.choice(WEIRD_F)))
                elif ctype is C.c_uint32:
                    setattr(fp, name, int(rng.choice(WEIRD_U)) if name not in ("width", "height") else int(rng.choice([0, 1, 64, 65535, 70000, 2**31])))
                else:  # float arrays
                    arr = getattr(fp, name)
                    for i in range(len(arr)):
                        if rng.random() < 0.3:
                            arr[i] = float(rng.choice(WEIRD_F))
        try:
            if int(fp.width) * int(fp.height) > 4_000_000:
                raise g.VkgsError(-1, "skipped: frame too large for the probe")
            out = r.render(fp)
            frames += 1
        except g.VkgsError:
            errors += 1
        except AssertionError:
            errors += 1
    # the context still renders the reference frame bit for bit
    r.upload(s, g.default_options(front_to_back=1))
    ok = np.array_equal(r.render(good_fp)[0], want)
    log(f"{trials} trials: {frames} frames, {errors} rejected calls, context intact afterwards: {ok}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(run(int(sys.argv[1]) if len(sys.argv) > 1 else 300, g.GaussianSplatting(0)))
