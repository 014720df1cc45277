"""Benchmark log lines in the reference's format, so the reference's own `benchmark.py` can parse and
plot results of this path next to its Vulkan pipelines.

Formats restated from the reference's emitters:
  * sequence header  `ParameterSequence %d "%s" = {\\n%s\\n%s}`  (nvpro_core2/nvutils/parameter_sequencer.cpp:96)
  * timer lines      `Timeline "{}"; level {}; Timer "{}"; GPU; avg {}; min {}; max {}; last {}; CPU; avg ...; samples {};`
                     in integer microseconds (nvpro_core2/nvutils/profiler.cpp:55)
  * memory block     `BENCHMARK_ADV <id> { Memory Scene; Host used \\t..; Device Used \\t..; Device Allocated \\t..; (bytes) ... }`
                     (src/gaussian_splatting.cpp:2601-2617)
and consumed by benchmark.py:19-76 (`parse_benchmark`). Stage names are the reference's profiler
sections: "GPU Dist", "GPU Sort", "Rasterization" (src/gaussian_splatting.cpp:1324,1346,567).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence


@dataclass
class TimerStat:
    """GPU / CPU times of one profiler section over the averaged frames, in microseconds."""
    gpu_us: List[float] = field(default_factory=list)
    cpu_us: List[float] = field(default_factory=list)

    def add(self, gpu_us: float, cpu_us: float = 0.0):
        self.gpu_us.append(gpu_us)
        self.cpu_us.append(cpu_us)

    @staticmethod
    def _agg(v: Sequence[float]):
        if not v:
            return 0, 0, 0, 0
        return int(round(sum(v) / len(v))), int(round(min(v))), int(round(max(v))), int(round(v[-1]))

    def line(self, name: str, timeline: str = "Frame", level: int = 1) -> str:
        ga, gmin, gmax, gl = self._agg(self.gpu_us)
        ca, cmin, cmax, cl = self._agg(self.cpu_us)
        return (f'Timeline "{timeline}"; level {level}; Timer "{name}"; GPU; avg {ga}; min {gmin}; max {gmax}; last {gl}; '
                f'CPU; avg {ca}; min {cmin}; max {cmax}; last {cl}; samples {len(self.gpu_us)};')


def sequence_block(index: int, name: str, parameters: str, timers: Dict[str, TimerStat]) -> str:
    body = "\n".join(t.line(n) for n, t in timers.items())
    return f'ParameterSequence {index} "{name}" = {{\n{parameters}\n{body}\n}}\n'


def memory_block(index: int, scene=(0, 0, 0), rasterization=(0, 0, 0), raytracing=(0, 0, 0)) -> str:
    def row(tag, m):
        return f" Memory {tag}; Host used \t{int(m[0])}; Device Used \t{int(m[1])}; Device Allocated \t{int(m[2])}; (bytes)"
    return "\n".join([f"BENCHMARK_ADV {index} {{", row("Scene", scene), row("Rasterization", rasterization),
                      row("Raytracing", raytracing), "}"]) + "\n"


def run_sequence(renderer, fp, index: int, name: str, parameters: str = "", frames: int = 1024, averages: int = 128) -> str:
    """Render `frames` frames and report the last `averages` of them (the reference's --sequenceframes /
    --sequenceaverages, benchmark_3dgs.cfg:2-3): per-stage device times from the context's own
    per-kernel events, one frame at a time like the reference's profiler sections."""
    import time
    renderer.set_frames_in_flight(1)
    for _ in range(max(0, frames - averages)):
        renderer.render_async(fp)
    renderer.sync()
    renderer.set_profiling(True)
    timers = {"GPU Dist": TimerStat(), "GPU Sort": TimerStat(), "Rasterization": TimerStat(), "Frame": TimerStat()}
    for _ in range(averages):
        t0 = time.perf_counter()
        renderer.render_async(fp)
        cpu_us = (time.perf_counter() - t0) * 1e6
        st = renderer.last_frame_stats()
        timers["GPU Dist"].add(st.ms_dist * 1e3)
        timers["GPU Sort"].add(st.ms_sort * 1e3)
        timers["Rasterization"].add(st.ms_raster * 1e3)
        timers["Frame"].add(st.ms_total * 1e3, cpu_us)
    renderer.set_profiling(False)
    renderer.set_frames_in_flight(2)
    return sequence_block(index, name, parameters, timers)
