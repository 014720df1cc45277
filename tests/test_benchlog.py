"""The benchmark log this repo prints is parsed by the reference's own benchmark.py regexes
(benchmark.py:19-23, restated here) — timers in integer microseconds, memory block per sequence."""
import re

from vk_gaussian_splatting_b200 import benchlog as B

# the four patterns of parse_benchmark() in the reference's benchmark.py:19-23
BENCHMARK = re.compile(r'ParameterSequence\s+(\d+)\s+"([^"]+)"\s*=')
TIMER = re.compile(r'Timer\s+"([^"]+)"\s*;\s*GPU;\s*avg\s+(\d+);.*?CPU;\s*avg\s+(\d+);')
ADV = re.compile(r'BENCHMARK_ADV (\d+) {')
MEMORY = re.compile(r'Memory (\w+); Host used\s+(\d+); Device Used\s+(\d+); Device Allocated\s+(\d+);')


def test_log_round_trips_through_the_reference_parser():
    log = ""
    for i, name in enumerate(("CUDA pipeline fp32", "CUDA pipeline fp16")):
        t = {"GPU Dist": B.TimerStat(), "GPU Sort": B.TimerStat(), "Rasterization": B.TimerStat()}
        for k in range(4):
            t["GPU Dist"].add(78.4 + k, 3.0)
            t["GPU Sort"].add(66.0, 2.0)
            t["Rasterization"].add(250.6 - k, 9.6)
        log += B.sequence_block(i, name, "--shformat %d" % i, t) + B.memory_block(i, (100, 200, 300), (0, 400, 500))
    parts = re.split(BENCHMARK, log)[1:]
    assert len(parts) == 6 and [int(parts[0]), parts[1], int(parts[3]), parts[4]] == [0, "CUDA pipeline fp32", 1, "CUDA pipeline fp16"]
    timers = {m.group(1): (int(m.group(2)), int(m.group(3))) for m in TIMER.finditer(parts[2])}
    assert timers == {"GPU Dist": (80, 3), "GPU Sort": (66, 2), "Rasterization": (249, 10)}
    adv = re.split(ADV, log)[1:]
    assert [int(adv[0]), int(adv[2])] == [0, 1]
    mem = {m.group(1): tuple(int(m.group(k)) for k in (2, 3, 4)) for m in MEMORY.finditer(adv[1])}
    assert mem == {"Scene": (100, 200, 300), "Rasterization": (0, 400, 500), "Raytracing": (0, 0, 0)}
