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

Prints ONE JSON line (rank 0). `value` = frames/s with the scene resident in HBM and the frame left
in HBM; `e2e` = the same through the synchronous C-ABI call with host frame parameters in and the
fp32 RGBA frame copied to pinned host memory every step.
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


def cpu_sorter_baseline(scene, cam, budget_s=12.0, max_reps=20):
    """The reference's CPU sorting path (splat_sorter_async restatement in oracle/) on all host cores."""
    import numpy as np
    from oracle import oracle as O
    threads = O.hardware_concurrency()
    eye = np.array(cam.eye, np.float32)
    direction = np.array(cam.ctr, np.float32) - eye
    ident = np.eye(4, dtype=np.float32).reshape(16)
    best, t_start, reps = None, time.perf_counter(), 0
    O.cpu_sort(scene.positions, ident, direction, eye, front_to_back=True, mode=1, threads=threads)  # warm-up
    while reps < max_reps and (time.perf_counter() - t_start) < budget_s:
        _, _, ms_d, ms_s = O.cpu_sort(scene.positions, ident, direction, eye, front_to_back=True, mode=1, threads=threads)
        if best is None or ms_d + ms_s < best[0] + best[1]:
            best = (ms_d, ms_s)
        reps += 1
    return {"ms_dist": best[0], "ms_sort": best[1], "cores": threads, "reps": reps}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (its CPU sorting mode:
    SplatSorterAsync::innerSort, all host threads). The reference has no CPU rasterizer, so a
    reference 'frame' is CPU Dist + CPU Sort of all N splats; rank 0 alone runs it."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    import numpy as np
    import vk_gaussian_splatting_b200 as g
    from oracle import oracle as O
    scene = g.synth_scene(N_SPLATS, SH_DEGREE, SEED)
    cam = g.default_camera()
    threads = O.hardware_concurrency()
    eye = np.array(cam.eye, np.float32)
    direction = np.array(cam.ctr, np.float32) - eye
    ident = np.eye(4, dtype=np.float32).reshape(16)
    for _ in range(args.warmup):
        O.cpu_sort(scene.positions, ident, direction, eye, True, 1, threads)
    t0 = time.perf_counter()
    dsum = ssum = 0.0
    for _ in range(args.steps):
        _, _, md, ms = O.cpu_sort(scene.positions, ident, direction, eye, True, 1, threads)
        dsum += md
        ssum += ms
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {
        "impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "splats": N_SPLATS, "note": "CPU Dist + CPU Sort only; the reference has no CPU rasterizer"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"full sort of all {N_SPLATS} splats per step (splat_sorter_async restatement, "
                                   f"__gnu_parallel::sort on {threads} threads); ms_dist={dsum / args.steps:.2f} ms_sort={ssum / args.steps:.2f}"},
        "msplats_per_sec": fps * N_SPLATS / 1e6,
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    farm = F.Farm(backend="nccl", device="cuda")  # no-op control plane when world == 1
    barrier, max_over_ranks = farm.barrier, farm.max_over_ranks

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

    # ---- device-resident throughput ---------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(args.warmup):
        r.render_async(fp)
    r.sync()
    barrier()
    t_begin = time.time()
    l0 = r.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        r.render_async(fp)
    e1.record(stream)
    torch.cuda.synchronize()
    t_end = time.time()
    launches = r.launch_count() - l0
    r.sync()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    st = r.last_frame_stats()
    fps = world * args.steps / (ms_total / 1000.0)

    # ---- end to end: host params in, RGBA frame to pinned host memory out, synchronous call ---------
    def e2e_run(target_fmt, torch_dtype):
        r.set_target_format(target_fmt)
        host_imgs = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch_dtype, pin_memory=True) for _ in range(2)]
        host_np = [t.numpy() for t in host_imgs]
        steps = max(10, min(args.steps, 100))
        for i in range(3):
            r.render_to_host_async(fp, host_np[i % 2])
        r.sync()
        barrier()
        e0.record(stream)
        for i in range(steps):
            # the call a user makes: host frame parameters in, finished RGBA frame copied to pinned host
            # memory out, every step; several frames in flight so step i's copy overlaps step i+1's kernels
            r.render_to_host_async(fp, host_np[i % 2])
        e1.record(stream)
        r.sync()
        torch.cuda.synchronize()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        return world * steps / (ms / 1000.0), ms / steps, steps, host_imgs[0].element_size() * WIDTH * HEIGHT * 4

    # headline e2e: the reference's default colour target (COLOR_MAIN = R16G16B16A16_SFLOAT); fp32 target alongside
    fps_e2e, ms_e2e_step, e2e_steps, d2h_bytes = e2e_run(A.FORMAT_FLOAT16, torch.float16)
    fps_e2e32, ms_e2e32_step, _, d2h_bytes32 = e2e_run(A.FORMAT_FLOAT32, torch.float32)
    fps_e2e8, ms_e2e8_step, _, d2h_bytes8 = e2e_run(A.FORMAT_UINT8, torch.uint8)
    r.set_target_format(A.FORMAT_FLOAT32)

    # ---- per-kernel profile (separate frames, cudaEvents around every launch on the launch stream) --
    r.set_frames_in_flight(1)  # per-kernel times need one frame at a time (no cross-frame overlap)
    r.set_profiling(True)
    acc, nprof = {}, 10
    for _ in range(nprof):
        for _ in range(8):  # steady state: frames back to back, events of the last one are read
            r.render_async(fp)
        s = r.last_frame_stats()
        for k, v in s.ms_kernel.items():
            acc[k] = acc.get(k, 0.0) + v / nprof
    r.set_profiling(False)
    # blend workload of this frame (separate, untimed frame with the counting variant of the blend kernel)
    copt = g.default_options(front_to_back=1, transmittance_epsilon=EPS)
    copt._reserved[0] = 128
    r.upload(scene, copt)
    r.render_async(fp)
    cst = r.last_frame_stats()
    r.upload(scene, opt)
    r.set_frames_in_flight(4)
    stage_ms = {"GPU Dist": acc["preprocess"], "GPU Sort": acc["sort_hist"] + sum(acc[f"sort_pass{i}"] for i in range(4)),
                "Rasterization": acc["bin_emit"] + acc["tile_hist"] + acc["tile_sort0"] + acc["tile_sort1"] + acc["tile_ranges"] + acc["blend"]}
    dominant = max(acc, key=acc.get)
    n, v, p, d = N_SPLATS, st.visible_count, WIDTH * HEIGHT, st.tile_pairs
    # algorithmic bytes per launch (SURVEY.md §8d per-unit figures; DESIGN.md "Kernels")
    alg = {"preprocess": 12 * n + 8 * v + 236 * v, "sort_hist": 4 * v, "bin_emit": 4 * v + 8 * v, "tile_hist": 4 * d,
           "tile_ranges": 4 * d, "blend": 16 * p, "tile_sort0": 16 * d, "tile_sort1": 16 * d}
    for i in range(4):
        alg[f"sort_pass{i}"] = 16 * v
    peak, peak_src = peaks()
    ach = alg[dominant] / (acc[dominant] * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(dominant), "traffic_source": traffic.get("_source"), "peak_source": peak_src,
                "kernel_ms": acc[dominant],
                "note": ("k_blend is bounded by fp32/SFU instruction issue, not HBM (SURVEY.md 8d): its GB/s fraction is low by "
                         "nature; the HBM-bound kernel of the frame is reported under hbm_kernel") if dominant == "blend" else None,
                "per_kernel_gbs": {k: alg[k] / (acc[k] * 1e-3) / 1e9 for k in alg if acc.get(k, 0) > 0}}
    # the kernel that streams the splat attributes (91 % of the frame's algorithmic bytes) against the same peak
    pre_ach = alg["preprocess"] / (acc["preprocess"] * 1e-3) / 1e9
    roofline["hbm_kernel"] = {"kernel": "preprocess", "achieved": pre_ach, "frac": pre_ach / peak, "kernel_ms": acc["preprocess"],
                              "algorithmic_bytes": alg["preprocess"], "traffic": traffic.get("preprocess")}
    whole = st.bytes_algorithmic / (ms_total / args.steps * 1e-3) / 1e9

    line = None
    if rank == 0:
        line = {
            "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "splats": N_SPLATS, "sh_degree": SH_DEGREE, "width": WIDTH, "height": HEIGHT,
                       "views": "one camera per GPU (rank r: default eye rotated r*45deg about +Y)", "seed": hex(SEED),
                       "transmittance_epsilon": EPS,
                       "l2": "per-frame inputs (232 MB of splat attributes) exceed the 126 MB L2; no explicit flush"},
            "msplats_per_sec": fps * N_SPLATS / 1e6, "visible_splats": v, "tile_pairs": d,
            "stage_ms": stage_ms, "kernel_ms": acc,
            "frame_algorithmic_bytes": st.bytes_algorithmic, "frame_hbm_gbs": whole, "frame_hbm_frac": whole / peak,
            "blend_workload": {"list_entries_evaluated_x64px": cst.list_entries_evaluated, "fragments_blended": cst.fragments_blended,
                               "fragments_per_s": cst.fragments_blended / (acc["blend"] * 1e-3),
                               "pixel_evaluations_per_s": 64 * cst.list_entries_evaluated / (acc["blend"] * 1e-3),
                               "note": "the blend stage is bounded by fp32 issue: its natural unit is fragments/s (BASELINE.md 3)"},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": C.sizeof(A.FrameParams),
                    "d2h_bytes_per_step": d2h_bytes + 48, "steps": e2e_steps, "ms_per_step": ms_e2e_step,
                    "target": "RGBA16F (reference default COLOR_MAIN format), pinned host buffer, 4 frames in flight",
                    "note": "PCIe-bound: d2h_bytes_per_step x value is the host link bandwidth; the other target formats show it",
                    "fp32_target": {"value": fps_e2e32, "ms_per_step": ms_e2e32_step, "d2h_bytes_per_step": d2h_bytes32 + 48},
                    "rgba8_target": {"value": fps_e2e8, "ms_per_step": ms_e2e8_step, "d2h_bytes_per_step": d2h_bytes8 + 48}},
        }
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_sorter_baseline(scene, cam)
            f = 1000.0 / (cb["ms_dist"] + cb["ms_sort"])
            line["cpu_baseline"] = {"value": f, "unit": "frames/s", "cores": cb["cores"], "kind": "port",
                                    "sample": f"best of {cb['reps']} full sorts of all {N_SPLATS} splats (CPU Dist {cb['ms_dist']:.2f} ms + "
                                              f"CPU Sort {cb['ms_sort']:.2f} ms, splat_sorter_async restatement, __gnu_parallel::sort); "
                                              "raster excluded: the reference has no CPU rasterizer"}
            # the whole path (dist + sort + projection + raster + blend) once through the CPU oracle, same workload
            try:
                from oracle import oracle as O
                t0 = time.perf_counter()
                pk = O.Packed(scene)
                O.render(pk, O.frame_params(cam, WIDTH, HEIGHT), O.default_options(front_to_back=1))
                line["cpu_baseline"]["oracle_frame"] = {"ms": 1000.0 * (time.perf_counter() - t0), "cores": O.render_threads(),
                                                        "sample": "one frame of the workload through oracle/vkgs_oracle.c (pack + dist + sort + "
                                                                  "per-splat + raster in row bands), exact compositing (no early termination)"}
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
