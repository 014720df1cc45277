"""Multi-view farm: the one place this path shards. Frames for different cameras are independent, so
every GPU (one process per GPU, torch.distributed over NCCL/NVLink) holds the whole splat set and
renders its own view; there is no data-path collective. torch.distributed only carries the start /
stop barriers and the reduction of the per-rank device timings (max over ranks) — SURVEY.md §8(e).

Host-side logic only (CPU-testable with the gloo backend); the rendering itself is api.GaussianSplatting.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

from . import _abi as A
from .api import orbit_camera


@dataclass
class RankInfo:
    rank: int
    local_rank: int
    world: int


def rank_info() -> RankInfo:
    """RANK / LOCAL_RANK / WORLD_SIZE as set by torch.distributed.run (defaults: single process)."""
    return RankInfo(int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))


def _gpu_cpu_affinity(local_rank: int):
    """CPUs next to the GPU as NVML reports them (ideal CPU affinity), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        return [c for c in cpus if c < (os.cpu_count() or 0)] or None
    except Exception:
        return None


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Pin this process to the CPUs next to its GPU before any pinned host buffer is allocated (first touch puts the pages
    on that memory node): with eight ranks copying frames to the host at once, buffers on the wrong node all cross the
    socket interconnect. No-op when the topology is unknown."""
    cpus = _gpu_cpu_affinity(local_rank)
    if not cpus:
        return False
    try:
        allowed = os.sched_getaffinity(0)
        want = set(cpus) & allowed
        if want:
            os.sched_setaffinity(0, want)
            return True
    except Exception:
        pass
    return False


def numa_report(local_rank: int) -> dict:
    cpus = _gpu_cpu_affinity(local_rank)
    try:
        now = sorted(os.sched_getaffinity(0))
    except Exception:
        now = []
    def span(c):
        return f"{c[0]}-{c[-1]} ({len(c)})" if c else None
    return {"gpu_cpu_affinity": span(cpus) if cpus else None, "process_cpus": span(now)}


def view_for_rank(rank: int, n_views: int = 8) -> A.Camera:
    """View v -> rank v mod G. Rank 0 renders the reference's default camera; rank r the default eye
    rotated by r * 360/n_views degrees about +Y (8 eyes on a circle looking at the origin)."""
    return orbit_camera(rank % n_views, n_views)


def views_for_rank(rank: int, world: int, n_views: int) -> List[int]:
    """All view indices a rank owns when there are more views than GPUs (round-robin)."""
    return [v for v in range(n_views) if v % world == rank]


class Farm:
    """Thin wrapper over torch.distributed for the farm's control plane."""

    def __init__(self, backend: Optional[str] = None, device: Optional[str] = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.info = rank_info()
        self.device = device or ("cuda" if backend == "nccl" else "cpu")
        self.active = self.info.world > 1
        if self.active and not dist.is_initialized():
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", self.info.local_rank)
            # NCCL prints its version banner on STDOUT when the first communicator comes up (NCCL_DEBUG=VERSION, the
            # default of some images); callers such as bench.py own stdout (one JSON line), so the file descriptor is
            # pointed at stderr while the communicator is created and the first collective runs
            import sys
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group(backend or "gloo", **kw)
                dist.barrier()
                if backend == "nccl":
                    torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)

    def barrier(self):
        if self.active:
            self.dist.barrier()
        if self.device == "cuda":
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if not self.active:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if not self.active:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_timings(self, local: dict) -> Optional[list]:
        """Fixed-size per-rank timing record gathered on every rank (returned on rank 0 only)."""
        if not self.active:
            return [local]
        keys = sorted(local)
        t = self.torch.tensor([float(local[k]) for k in keys], dtype=self.torch.float64, device=self.device)
        out = [self.torch.zeros_like(t) for _ in range(self.info.world)]
        self.dist.all_gather(out, t)
        if self.info.rank != 0:
            return None
        return [dict(zip(keys, o.tolist())) for o in out]

    def aggregate_fps(self, frames_per_rank: int, ms_local: float) -> float:
        """Whole-job frames/s: all ranks' frames over the slowest rank's device time."""
        ms = self.max_over_ranks(ms_local)
        return self.info.world * frames_per_rank / (ms / 1000.0)

    def close(self):
        if self.active and self.dist.is_initialized():
            self.dist.barrier()
            self.dist.destroy_process_group()
