#!/usr/bin/env python
"""bench.py — frames/s of the 3DGS forward raster path (dist/cull -> radix sort -> tile raster).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU sort path, host cores

A "step" is one frame of the hot path. Workload at every N: BASELINE.json configs[1] — 1 M-splat
synthetic scene, SH degree 3, 1920x1080, reference default camera, front-to-back compositing
(north_star). Multi-GPU (N>1, launched by torch.distributed.run, one rank per GPU) is the
embarrassingly parallel multi-view case: every rank holds the full scene and renders its own
camera (rank r = default eye rotated r*45 deg about +Y); no data-path collective, NCCL only
carries the barrier and the max-over-ranks timing ("weak" scaling: work per GPU is fixed).

Prints ONE JSON line (rank 0). `value` = frames/s with the scene resident in HBM and the frame left in HBM (four frames
in flight, like the reference's swapchain loop). `e2e` = the same through the asynchronous C-ABI call a streaming caller makes
(vkgs_render_to_host_async: host frame parameters in, RGBA16F frame — the reference's COLOR_MAIN format — copied to pinned host
memory every step); `e2e.sync_call` = the plain synchronous vkgs_render into pinned memory, one frame at a time (the call
INTEGRATION.md's binding makes). At N = 1 the line also carries the other BASELINE.json configurations (`configs`: 6 M / 4K,
30 M / 1080p, the 30 M sort-only point), the reference-default back-to-front order (`btf`) and the CPU baselines; at N > 1 the
multi-view farm of configs[3] (`farm_cfg4`: 6 M splats, one view per GPU), a cross-rank frame check and a host-link probe.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_SPLATS = 1_000_000
SH_DEGREE = 3
WIDTH, HEIGHT = 1920, 1080
SEED = 0x3D650001  # 0x3D650000 + config index
EPS = 2.0 ** -15   # front-to-back early-out; error bound eps*max|rgb| < 1e-4 (tests/test_gpu_parity.py)
WORKLOAD = "configs[1]: 1M-splat synthetic scene, SH degree 3, 1920x1080, default camera, front-to-back"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM bytes (read + write) of each kernel from the committed `ncu --set full` capture of one frame of this
    workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py). {} when the file is absent."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if not p.exists():
        return {}
    try:
        d = json.loads(p.read_text())
        k = d["kernels"]
        out = {"_source": "profiles/ncu_traffic.json (" + d.get("source", "?") + ")"}
        if "k_preprocess" in k:
            out["preprocess"] = sum(x["dram_bytes"] for x in k["k_preprocess"])
        if "k_blend" in k:
            out["blend"] = k["k_blend"][0]["dram_bytes"]
        if "k_bin_emit" in k:
            out["bin_emit"] = k["k_bin_emit"][0]["dram_bytes"]
        return out
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (started before the warm-up;
    only samples whose timestamps fall inside the timed window are used)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []   # (host time of arrival, fields)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t_begin=None, t_end=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t_begin is None or (t_begin - 0.02 <= t <= t_end + 0.04)]
        window = "timed region"
        if not rows:
            rows, window = [r for (_, r) in self.rows], "whole run (timed region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


def cpu_sorter_baseline(positions, budget_s=10.0, max_reps=20):
    """The reference's CPU sorting path (splat_sorter_async restatement in oracle/) on all host cores."""
    import numpy as np
    from oracle import oracle as O
    cam = O.default_camera()
    threads = O.hardware_concurrency()
    eye = np.array(cam.eye, np.float32)
    direction = np.array(cam.ctr, np.float32) - eye
    ident = np.eye(4, dtype=np.float32).reshape(16)
    best, t_start, reps = None, time.perf_counter(), 0
    O.cpu_sort(positions, ident, direction, eye, front_to_back=True, mode=1, threads=threads)  # warm-up
    while reps < max_reps and (time.perf_counter() - t_start) < budget_s:
        _, _, ms_d, ms_s = O.cpu_sort(positions, ident, direction, eye, front_to_back=True, mode=1, threads=threads)
        if best is None or ms_d + ms_s < best[0] + best[1]:
            best = (ms_d, ms_s)
        reps += 1
    return {"ms_dist": best[0], "ms_sort": best[1], "cores": threads, "reps": reps}


def config0_cpu_plumbing():
    """BASELINE configs[0], CPU only, once: 100 k Gaussians, SH0, 512x512, CPU sorter order -> CPU blend in that order."""
    import numpy as np
    import vk_gaussian_splatting_b200 as g
    from oracle import oracle as O
    s = g.synth_scene(100_000, 0, 0x3D650000)
    cam = O.default_camera()
    eye = np.array(cam.eye, np.float32)
    ident = np.eye(4, dtype=np.float32).reshape(16)
    t0 = time.perf_counter()
    order, _, ms_d, ms_s = O.cpu_sort(s.positions, ident, np.array(cam.ctr, np.float32) - eye, eye, front_to_back=False, mode=1)
    t1 = time.perf_counter()
    pk = O.Packed(s)
    img = O.render_presorted(pk, O.frame_params(cam, 512, 512), O.default_options(front_to_back=0), order)
    t2 = time.perf_counter()
    return {"workload": "configs[0]: 100k Gaussians, SH0, 512x512, CPU splat_sorter_async order + CPU blend (back-to-front)",
            "ms_cpu_dist": ms_d, "ms_cpu_sort": ms_s, "ms_pack_project_blend": 1000.0 * (t2 - t1), "fps": 1.0 / (t2 - t0),
            "threads": O.render_threads(), "alpha_max": float(img[..., 3].max())}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its CPU sorting mode:
    SplatSorterAsync::innerSort, all host threads). The reference has no CPU rasterizer, so a
    reference 'frame' is CPU Dist + CPU Sort of all N splats; rank 0 alone runs it. Nothing of the product is loaded here:
    the scene comes from the oracle's numpy restatement of the generator (bit-identical, tests/test_config0_cpu.py)."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import numpy as np
    from oracle import oracle as O
    positions = O.synth_positions(N_SPLATS, SEED)
    cam = O.default_camera()
    threads = O.hardware_concurrency()
    eye = np.array(cam.eye, np.float32)
    direction = np.array(cam.ctr, np.float32) - eye
    ident = np.eye(4, dtype=np.float32).reshape(16)
    for _ in range(args.warmup):
        O.cpu_sort(positions, ident, direction, eye, True, 1, threads)
    steps = args.steps  # one step = one full CPU Dist + CPU Sort of the 1 M splats (~10 ms on 16+ cores)
    t0 = time.perf_counter()
    dsum = ssum = 0.0
    for _ in range(steps):
        _, _, md, ms = O.cpu_sort(positions, ident, direction, eye, True, 1, threads)
        dsum += md
        ssum += ms
    dt = time.perf_counter() - t0
    fps = steps / dt
    line = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "splats": N_SPLATS, "note": "CPU Dist + CPU Sort only; the reference has no CPU rasterizer"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"full sort of all {N_SPLATS} splats per step (splat_sorter_async restatement, "
                                   f"__gnu_parallel::sort on {threads} threads); ms_dist={dsum / steps:.2f} ms_sort={ssum / steps:.2f}"},
        "msplats_per_sec": fps * N_SPLATS / 1e6,
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def frame_hash(a) -> int:
    """Order-dependent 64-bit checksum of a frame's bits (cross-rank frame check)."""
    import numpy as np
    v = np.ascontiguousarray(a).view(np.uint32).astype(np.uint64).ravel()
    w = (np.arange(v.size, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return int((v * w).sum(dtype=np.uint64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3 / cfg5 / sort-only / btf blocks (quick runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import vk_gaussian_splatting_b200 as g
    from vk_gaussian_splatting_b200 import _abi as A

    from vk_gaussian_splatting_b200 import farm as F

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    rank, local, world = dist_env()
    torch.cuda.set_device(local)
    # keep stdout to the ONE JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    F.bind_to_gpu_numa_node(local)  # pinned host buffers land on the memory node next to this rank's GPU
    farm = F.Farm(backend="nccl", device="cuda")  # no-op control plane when world == 1
    barrier, max_over_ranks = farm.barrier, farm.max_over_ranks
    peak, peak_src = peaks()

    scene = g.synth_scene(N_SPLATS, SH_DEGREE, SEED)  # every rank regenerates the scene from the seed
    cam = F.view_for_rank(rank, 8)                    # rank 0 = the reference default camera
    fp = g.frame_params(cam, WIDTH, HEIGHT)
    opt = g.default_options(front_to_back=1, transmittance_epsilon=EPS)

    # a real (non-default) torch stream: the context launches on it, torch.cuda.Event times it
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    r = g.GaussianSplatting(local, stream=stream.cuda_stream)
    r.upload(scene, opt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_frames(frame_params, steps, warmup, segments=1):
        """`steps` frames back to back (frames in flight as configured), CUDA events on the caller's stream, max over ranks.
        Returns (ms_total, [ms of each segment])."""
        for _ in range(warmup):
            r.render_async(frame_params)
        r.sync()
        barrier()
        timed_frames.launches0 = r.launch_count()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(segments + 1)]
        bounds = [steps * k // segments for k in range(segments + 1)]
        marks[0].record(stream)
        k = 1
        for i in range(steps):
            r.render_async(frame_params)
            if i + 1 == bounds[k]:
                marks[k].record(stream)
                k += 1
        torch.cuda.synchronize()
        timed_frames.launches = r.launch_count() - timed_frames.launches0  # kernels of this context launched inside the timed region
        r.sync()
        barrier()
        return max_over_ranks(marks[0].elapsed_time(marks[-1])), [marks[j].elapsed_time(marks[j + 1]) for j in range(segments)]

    def kernel_profile(frame_params, frames_in_flight, nprof=10):
        """Per-kernel device times (cudaEvents around every launch): `frames_in_flight` = 1 times each kernel alone, 4 times
        it the way `value` runs it (thin front end co-running with other frames' kernels: inflated by the overlap)."""
        r.set_frames_in_flight(frames_in_flight)
        r.set_profiling(True)
        acc = {}
        for _ in range(nprof):
            for _ in range(8):  # steady state: frames back to back, events of the last one are read
                r.render_async(frame_params)
            s = r.last_frame_stats()
            for k, v in s.ms_kernel.items():
                acc[k] = acc.get(k, 0.0) + v / nprof
        r.set_profiling(False)
        r.set_frames_in_flight(4)
        return acc, s

    def stage_ms_of(acc):
        return {"GPU Dist": acc["preprocess"], "GPU Sort": acc["sort_hist"] + sum(acc[f"sort_pass{i}"] for i in range(4)),
                "Rasterization": acc["bin_emit"] + acc["tile_hist"] + acc["tile_sort0"] + acc["tile_sort1"] + acc["tile_ranges"] + acc["blend"]}

    # ---- device-resident throughput ---------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    # Untimed pre-warm on top of the W warm-up frames: a few hundred ms of frames so that the timed region (4 ms at the
    # driver's --steps 20) starts at boost clocks with every kernel loaded — measured 4240 vs 4535 frames/s at
    # --steps 20 without / with it; --steps 1000 is unaffected.
    PREWARM_S = 0.3
    t_pw, prewarm_frames = time.time(), 0
    while time.time() - t_pw < PREWARM_S:
        for _ in range(16):
            r.render_async(fp)
        r.sync()
        prewarm_frames += 16
    t_begin = time.time()
    ms_total, seg_ms = timed_frames(fp, args.steps, args.warmup, segments=min(10, args.steps))
    t_end = time.time()
    launches = timed_frames.launches
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    st = r.last_frame_stats()
    fps = world * args.steps / (ms_total / 1000.0)
    seg_steps = [args.steps * (k + 1) // len(seg_ms) - args.steps * k // len(seg_ms) for k in range(len(seg_ms))]
    seg_per_step = [m / n for m, n in zip(seg_ms, seg_steps) if n]

    # ---- end to end: host params in, RGBA frame to pinned host memory out ------------------------------------------
    def e2e_run(target_fmt, torch_dtype, sync_call=False):
        r.set_target_format(target_fmt)
        host_imgs = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch_dtype, pin_memory=True) for _ in range(4)]
        host_np = [t.numpy() for t in host_imgs]
        # its own step count (reported as e2e.steps): at least 100 so that the fill and drain of the four-deep pipeline
        # (about one frame latency, 0.6 ms) do not dominate a 20-step region
        steps = max(100, min(args.steps, 400))
        # three repetitions of the timed region, the median is reported (the copy rides on the host's memory system:
        # a single 30 ms region now and then catches a host-side hiccup); all three are listed in `e2e_repeats`
        reps = []
        for _rep in range(3):
            if sync_call:
                # the plain synchronous entry point: one frame at a time, returns when the frame is in host memory
                r.set_frames_in_flight(1)
                for i in range(3):
                    r.render(fp, out=host_np[0])
                barrier()
                t0 = time.perf_counter()
                for i in range(steps):
                    r.render(fp, out=host_np[i % 4])
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1000.0
                r.set_frames_in_flight(4)
                barrier()
                reps.append(max_over_ranks(wall))  # host clock: the call blocks, so the caller's wall time IS the latency
            else:
                for i in range(4):
                    r.render_to_host_async(fp, host_np[i % 4])
                r.sync()
                barrier()
                e0.record(stream)
                for i in range(steps):
                    # the call a streaming caller makes: host frame parameters in, finished RGBA frame copied to pinned host
                    # memory out, every step; four frames in flight so step i's copy overlaps step i+1's kernels
                    r.render_to_host_async(fp, host_np[i % 4])
                e1.record(stream)
                r.sync()
                torch.cuda.synchronize()
                barrier()
                reps.append(max_over_ranks(e0.elapsed_time(e1)))
        ms = statistics.median(reps)
        e2e_run.repeats_ms_per_step = [x / steps for x in reps]
        return world * steps / (ms / 1000.0), ms / steps, steps, host_imgs[0].element_size() * WIDTH * HEIGHT * 4, host_np[0]

    # headline e2e: the reference's default colour target (COLOR_MAIN = R16G16B16A16_SFLOAT); the other targets alongside
    fps_e2e, ms_e2e_step, e2e_steps, d2h_bytes, frame16 = e2e_run(A.FORMAT_FLOAT16, torch.float16)
    e2e_repeats = list(e2e_run.repeats_ms_per_step)
    my_hash = frame_hash(frame16)
    fps_sync, ms_sync_step, _, _, _ = e2e_run(A.FORMAT_FLOAT16, torch.float16, sync_call=True)
    sync_repeats = list(e2e_run.repeats_ms_per_step)
    fps_e2e32, ms_e2e32_step, _, d2h_bytes32, _ = e2e_run(A.FORMAT_FLOAT32, torch.float32)
    fps_e2e8, ms_e2e8_step, _, d2h_bytes8, _ = e2e_run(A.FORMAT_UINT8, torch.uint8)
    r.set_target_format(A.FORMAT_FLOAT32)

    # ---- host link probe: what the e2e figure is bounded by (all ranks copy at the same time) ------------------
    def host_link_probe(nbytes=64 << 20, reps=8):
        dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        out = {}
        for name, (dst, src) in (("d2h", (host, dev)), ("h2d", (dev, host))):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            barrier()
            e0.record(stream)
            for _ in range(reps):
                dst.copy_(src, non_blocking=True)
            e1.record(stream)
            torch.cuda.synchronize()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            out[name + "_gbs_per_gpu"] = nbytes * reps / (ms * 1e-3) / 1e9
            out[name + "_gbs_aggregate"] = world * nbytes * reps / (ms * 1e-3) / 1e9
        return out
    link = host_link_probe()
    link["numa"] = F.numa_report(local)
    # how much of the measured link the end-to-end figure uses (per GPU, all ranks copying at the same time)
    link["e2e_d2h_gbs_per_gpu"] = d2h_bytes * (fps_e2e / world) / 1e9
    link["e2e_frac_of_link"] = link["e2e_d2h_gbs_per_gpu"] / link["d2h_gbs_per_gpu"]
    link["note"] = ("e2e is bounded by the host link: every step copies one RGBA16F frame to pinned host memory; with N ranks copying at "
                    "once the box's aggregate device-to-host bandwidth is shared (d2h_gbs_aggregate), so e2e stops scaling where the "
                    "device-resident `value` does not")

    # ---- cross-rank frame check: rank r's RGBA16F frame == view r rendered on rank 0 -----------------------------
    farm_check = None
    if world > 1:
        hashes = farm.gather_timings({"h_lo": float(my_hash & 0xffffff), "h_mid": float((my_hash >> 24) & 0xffffff), "h_hi": float(my_hash >> 48)})
        if rank == 0:
            r.set_target_format(A.FORMAT_FLOAT16)
            ok = []
            buf = np.empty((HEIGHT, WIDTH, 4), np.float16)
            for v in range(world):
                r.render(g.frame_params(F.view_for_rank(v, 8), WIDTH, HEIGHT), out=buf)
                h = frame_hash(buf)
                got = int(hashes[v]["h_lo"]) | (int(hashes[v]["h_mid"]) << 24) | (int(hashes[v]["h_hi"]) << 48)
                ok.append(got == h)
            r.set_target_format(A.FORMAT_FLOAT32)
            farm_check = {"views_match_rank0": all(ok), "per_view": ok,
                          "how": "64-bit checksum of every rank's RGBA16F frame vs the same view rendered on rank 0 (bit-identical frames)"}

    # ---- per-kernel profile: alone (1 frame in flight) and the way `value` runs (4 in flight) --------------------
    acc, _ = kernel_profile(fp, 1)
    acc4, _ = kernel_profile(fp, 4)
    # blend workload of this frame (separate, untimed frame with the counting variant of the blend kernel)
    copt = g.default_options(front_to_back=1, transmittance_epsilon=EPS)
    copt._reserved[0] = 128
    r.upload(scene, copt)
    r.render_async(fp)
    cst = r.last_frame_stats()
    r.upload(scene, opt)
    r.set_frames_in_flight(4)
    stage_ms = stage_ms_of(acc)
    dominant = max(acc, key=acc.get)
    n, v, p, d = N_SPLATS, st.visible_count, WIDTH * HEIGHT, st.tile_pairs

    def alg_bytes(n, v, p, d):
        # algorithmic bytes per launch (SURVEY.md §8d per-unit figures; DESIGN.md "Kernels")
        alg = {"preprocess": 12 * n + 8 * v + 236 * v, "sort_hist": 4 * v, "bin_emit": 4 * v + 8 * v, "tile_hist": 4 * d,
               "tile_ranges": 4 * d, "blend": 16 * p, "tile_sort0": 16 * d, "tile_sort1": 16 * d}
        for i in range(4):
            alg[f"sort_pass{i}"] = 16 * v
        return alg
    alg = alg_bytes(n, v, p, d)
    ach = alg[dominant] / (acc[dominant] * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(dominant), "traffic_source": traffic.get("_source"), "peak_source": peak_src,
                "kernel_ms": acc[dominant],
                "note": ("k_blend is bounded by fp32/SFU instruction issue, not HBM (SURVEY.md 8d): its GB/s fraction is low by "
                         "nature; the HBM-bound kernel of the frame is reported under hbm_kernel") if dominant == "blend" else None,
                "per_kernel_gbs": {k: alg[k] / (acc[k] * 1e-3) / 1e9 for k in alg if acc.get(k, 0) > 0}}
    # the kernel that streams the splat attributes (91 % of the frame's algorithmic bytes) against the same peak:
    # in the mode `value` is timed in (4 frames in flight, thin 1-CTA-per-SM launch beside other frames' kernels) and alone
    pre4 = alg["preprocess"] / (acc4["preprocess"] * 1e-3) / 1e9
    pre1 = alg["preprocess"] / (acc["preprocess"] * 1e-3) / 1e9
    roofline["hbm_kernel"] = {"kernel": "preprocess", "mode": "4 frames in flight (the mode `value` is timed in)", "achieved": pre4,
                              "frac": pre4 / peak, "kernel_ms": acc4["preprocess"], "algorithmic_bytes": alg["preprocess"],
                              "traffic": traffic.get("preprocess"),
                              "one_frame_in_flight": {"achieved": pre1, "frac": pre1 / peak, "kernel_ms": acc["preprocess"]}}
    whole = st.bytes_algorithmic / (ms_total / args.steps * 1e-3) / 1e9

    # ---- the other BASELINE.json configurations (rank 0 at N = 1; every rank for the farm at N > 1) -----------------
    def config_block(name, scene_n, w, h, seed, steps, warmup=5, **optkw):
        s = g.synth_scene(scene_n, SH_DEGREE, seed)
        r.upload(s, g.default_options(**optkw))
        fpc = g.frame_params(F.view_for_rank(rank, 8), w, h)
        r.set_frames_in_flight(4)
        ms, _ = timed_frames(fpc, steps, warmup)
        a1, stc = kernel_profile(fpc, 1, nprof=3)
        del s
        f = world * steps / (ms / 1000.0)
        balg = stc.bytes_algorithmic
        return {"workload": name, "splats": scene_n, "size": [w, h], "steps": steps, "fps": f, "ms_per_frame": ms / steps,
                "msplats_per_sec": f * scene_n / 1e6, "visible_splats": stc.visible_count, "tile_pairs": stc.tile_pairs,
                "stage_ms": stage_ms_of(a1), "kernel_ms": {k: v for k, v in a1.items() if v > 0.004},
                "frame_algorithmic_bytes": balg, "frame_hbm_gbs": balg / (ms / steps * 1e-3) / 1e9,
                "frame_hbm_frac": balg / (ms / steps * 1e-3) / 1e9 / peak}

    configs, btf, farm_cfg4 = None, None, None
    if not args.no_configs:
        ftb_kw = dict(front_to_back=1, transmittance_epsilon=EPS)
        if world == 1:
            configs = {
                "cfg3": config_block("configs[2]: 6M splats, SH3, 3840x2160, front-to-back", 6_000_000, 3840, 2160, 0x3D650002, 100, **ftb_kw),
                "cfg5": config_block("configs[4]: 30M splats, SH3, 1920x1080, front-to-back", 30_000_000, 1920, 1080, 0x3D650004, 30, **ftb_kw),
            }
            # sort-only point of configs[4]: 30 M random (key, value) pairs through the stand-alone sort (device time)
            rng = np.random.default_rng(5)
            m = 30_000_000
            keys = rng.integers(0, 1 << 32, size=m, dtype=np.uint64).astype(np.uint32)
            _, _, ms_sort = r.sort_pairs(keys, np.arange(m, dtype=np.uint32), repeats=5)
            del keys
            configs["sort_only_30m"] = {"pairs": m, "ms": ms_sort, "gpairs_per_sec": m / ms_sort / 1e6, "hbm_gbs_at_68B_per_pair": 68 * m / ms_sort / 1e6,
                                        "frac_of_peak": 68 * m / ms_sort / 1e6 / peak,
                                        "note": "32-bit keys, 4 passes + histogram; 68 B/pair is the onesweep minimum of SURVEY 8(d)"}
            # the reference's DEFAULT compositing order on the headline workload: back to front, additive alpha.
            # exact = no early termination (the reference's behaviour); eps = the same epsilon as the headline
            # (colour within eps*max|rgb|; the additive alpha then only sums the fragments composited before the stop)
            btf = {"exact": config_block("configs[1] back-to-front (reference default order), exact", N_SPLATS, WIDTH, HEIGHT, SEED, 200, front_to_back=0),
                   "eps": config_block("configs[1] back-to-front, transmittance_epsilon 2^-15", N_SPLATS, WIDTH, HEIGHT, SEED, 200,
                                       front_to_back=0, transmittance_epsilon=EPS),
                   "ftb_exact": config_block("configs[1] front-to-back, exact (no early termination)", N_SPLATS, WIDTH, HEIGHT, SEED, 200, front_to_back=1)}
            # the same frame through the VK3DGUT raster pipeline (unscented-transform projection, per-fragment
            # ray / particle response), reference defaults of that pipeline (EXTENT_CONIC, quadratic kernel, pinhole)
            configs["cfg2_3dgut"] = config_block("configs[1] through the VK3DGUT pipeline, front-to-back", N_SPLATS, WIDTH, HEIGHT, SEED, 100,
                                                 pipeline=A.PIPELINE_3DGUT, **ftb_kw)
        else:
            farm_cfg4 = config_block(f"configs[3]: 6M splats, SH3, 1920x1080, {world} independent views farmed over {world} GPUs (one view per GPU)",
                                     6_000_000, 1920, 1080, 0x3D650003, 100, **ftb_kw)
            farm_cfg4["views"] = world
        r.upload(scene, opt)

    line = None
    if rank == 0:
        line = {
            "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "splats": N_SPLATS, "sh_degree": SH_DEGREE, "width": WIDTH, "height": HEIGHT,
                       "views": "one camera per GPU (rank r: default eye rotated r*45deg about +Y)", "seed": hex(SEED),
                       "transmittance_epsilon": EPS, "frames_in_flight": 4,
                       "prewarm": f"{prewarm_frames} untimed frames ({PREWARM_S} s) before the {args.warmup} warm-up steps (clock ramp, lazy kernel loading)",
                       "l2": "per-frame inputs (232 MB of splat attributes) exceed the 126 MB L2; no explicit flush"},
            "ms_per_step_segments": {"n": len(seg_per_step), "min": min(seg_per_step), "max": max(seg_per_step),
                                     "mean": statistics.mean(seg_per_step), "stdev": statistics.pstdev(seg_per_step)},
            "msplats_per_sec": fps * N_SPLATS / 1e6, "visible_splats": v, "tile_pairs": d,
            "stage_ms": stage_ms, "kernel_ms": acc, "kernel_ms_4_in_flight": acc4,
            "frame_algorithmic_bytes": st.bytes_algorithmic, "frame_hbm_gbs": whole, "frame_hbm_frac": whole / peak,
            "blend_workload": {"list_entries_evaluated_x64px": cst.list_entries_evaluated, "fragments_blended": cst.fragments_blended,
                               "fragments_per_s": cst.fragments_blended / (acc["blend"] * 1e-3),
                               "pixel_evaluations_per_s": 64 * cst.list_entries_evaluated / (acc["blend"] * 1e-3),
                               "note": "the blend stage is bounded by fp32 issue: its natural unit is fragments/s (BASELINE.md 3)"},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": C.sizeof(A.FrameParams),
                    "d2h_bytes_per_step": d2h_bytes + 56, "steps": e2e_steps, "ms_per_step": ms_e2e_step,
                    "repeats_ms_per_step": e2e_repeats, "repeats": "median of three timed regions of `steps` steps each",
                    "target": "RGBA16F (reference default COLOR_MAIN format), pinned host buffer, 4 frames in flight (vkgs_render_to_host_async)",
                    "note": "bounded by the host link: d2h_bytes_per_step x value against host_link.d2h_gbs_per_gpu",
                    "sync_call": {"value": fps_sync, "ms_per_step": ms_sync_step, "repeats_ms_per_step": sync_repeats, "d2h_bytes_per_step": d2h_bytes + 56,
                                  "how": "plain vkgs_render (synchronous, one frame at a time) into pinned host memory, RGBA16F; host wall clock"},
                    "fp32_target": {"value": fps_e2e32, "ms_per_step": ms_e2e32_step, "d2h_bytes_per_step": d2h_bytes32 + 56},
                    "rgba8_target": {"value": fps_e2e8, "ms_per_step": ms_e2e8_step, "d2h_bytes_per_step": d2h_bytes8 + 56},
                    "host_link": link},
            "configs": configs, "btf": btf, "farm_cfg4": farm_cfg4, "farm_check": farm_check,
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle as O
            cb = cpu_sorter_baseline(scene.positions)
            f = 1000.0 / (cb["ms_dist"] + cb["ms_sort"])
            line["cpu_baseline"] = {"value": f, "unit": "frames/s", "cores": cb["cores"], "kind": "port",
                                    "sample": f"best of {cb['reps']} full sorts of all {N_SPLATS} splats (CPU Dist {cb['ms_dist']:.2f} ms + "
                                              f"CPU Sort {cb['ms_sort']:.2f} ms, splat_sorter_async restatement, __gnu_parallel::sort); "
                                              "raster excluded: the reference has no CPU rasterizer"}
            # the whole path (dist + sort + projection + raster + blend) once through the CPU oracle, same workload
            try:
                t0 = time.perf_counter()
                pk = O.Packed(scene)
                O.render(pk, O.frame_params(O.default_camera(), WIDTH, HEIGHT), O.default_options(front_to_back=1))
                line["cpu_baseline"]["oracle_frame"] = {"ms": 1000.0 * (time.perf_counter() - t0), "cores": O.render_threads(),
                                                        "sample": "one frame of the workload through oracle/vkgs_oracle.c (pack + dist + sort + "
                                                                  "per-splat + raster in row bands), exact compositing (no early termination)"}
                line["cpu_baseline"]["config0_cpu_plumbing"] = config0_cpu_plumbing()
            except Exception as e:  # the baseline must never take the bench line down
                line["cpu_baseline"]["oracle_frame"] = {"error": repr(e)}
        else:
            line["cpu_baseline"] = None
    r.close()
    farm.close()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
