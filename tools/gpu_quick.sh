#!/bin/bash
# Quick GPU check: parity tests + bench without the CPU baseline. usage: tools/gpu_quick.sh <tag> [pytest -k expr]
tag=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
python bench.py --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -4 gpurun_out/${tag}_tests.log
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("fps", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v * 1000, 1) for k, v in d["kernel_ms"].items() if v > 0.004})
PY
