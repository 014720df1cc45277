import sys, numpy as np
sys.path.insert(0,'/root/repo')
import vk_gaussian_splatting_b200 as g
s = g.synth_scene(1_000_000, 3, 0x3D650001)
fp = g.frame_params(g.default_camera(), 1920, 1080)
r = g.GaussianSplatting(0)
for ab in [0,1,2,4,8,3,7,15]:
    opt = g.default_options(front_to_back=1, transmittance_epsilon=2.0**-15)
    opt._reserved[0] = ab
    r.upload(s, opt)
    r.set_profiling(True)
    acc=[]
    for rep in range(5):
        for _ in range(6): r.render_async(fp)
        acc.append(r.last_frame_stats().ms_kernel['preprocess'])
    print('ablate', ab, 'preprocess ms', np.median(acc))
