import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import vk_gaussian_splatting_b200 as g
ab = int(sys.argv[1]) if len(sys.argv) > 1 else 0
s = g.synth_scene(1_000_000, 3, 0x3D650001)
fp = g.frame_params(g.default_camera(), 1920, 1080)
r = g.GaussianSplatting(0)
opt = g.default_options(front_to_back=1, transmittance_epsilon=2.0**-15)
opt._reserved[0] = ab
r.upload(s, opt)
for _ in range(4): r.render_async(fp)
r.sync()
