"""N>1 path on CPU: world_size-2 gloo run of the multi-view farm's control plane (view assignment,
barriers, max-over-ranks timing, timing gather). No GPU, no rendering."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from vk_gaussian_splatting_b200 import farm
    f = farm.Farm(backend="gloo", device="cpu")
    assert f.info.rank == rank and f.info.world == world and f.active
    cam = farm.view_for_rank(rank)
    f.barrier()
    ms_local = 10.0 + 5.0 * rank           # rank 1 is the slow one
    ms_max = f.max_over_ranks(ms_local)
    fps = f.aggregate_fps(frames_per_rank=100, ms_local=ms_local)
    rows = f.gather_timings({"ms": ms_local, "visible": 1000 + rank})
    total = f.sum_over_ranks(1.0)
    f.barrier()
    q.put((rank, list(cam.eye), ms_max, fps, rows, total, farm.views_for_rank(rank, world, 8)))
    f.close()


def test_world_size_2_gloo_farm():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, eye0, max0, fps0, rows0, tot0, views0), (r1, eye1, max1, fps1, rows1, tot1, views1) = res
    assert np.allclose(eye0, [1.7, 1.5, 1.7])               # rank 0 = reference default camera
    assert not np.allclose(eye0, eye1) and abs(eye1[1] - 1.5) < 1e-6
    assert abs(np.hypot(eye1[0], eye1[2]) - np.hypot(1.7, 1.7)) < 1e-5  # same orbit radius
    assert max0 == max1 == 15.0                              # max over ranks, identical everywhere
    assert fps0 == fps1 == 2 * 100 / 0.015                   # whole-job aggregate over the slowest rank
    assert rows1 is None and rows0 == [{"ms": 10.0, "visible": 1000.0}, {"ms": 15.0, "visible": 1001.0}]
    assert tot0 == tot1 == 2.0
    assert views0 == [0, 2, 4, 6] and views1 == [1, 3, 5, 7]


def test_single_process_farm_is_a_no_op():
    sys.path.insert(0, str(ROOT))
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    from vk_gaussian_splatting_b200 import farm
    f = farm.Farm(backend="gloo", device="cpu")
    assert not f.active and f.max_over_ranks(3.5) == 3.5 and f.gather_timings({"a": 1}) == [{"a": 1}]
    assert f.aggregate_fps(10, 5.0) == 10 / 0.005
    f.barrier()
    f.close()
