#!/bin/bash
# On the GPU box: the same bench_configs rows for this checkout (A) and for ab_wt/ (B, see tools/ab_prepare.sh).
mkdir -p gpurun_out
python tools/bench_configs.py "$@" > gpurun_out/ab_A.jsonl 2> gpurun_out/ab_A.err
(cd ab_wt && python tools/bench_configs.py "$@") > gpurun_out/ab_B.jsonl 2> gpurun_out/ab_B.err
python - <<'PY'
import json
for tag in "AB":
    for line in open(f"gpurun_out/ab_{tag}.jsonl"):
        try:
            d = json.loads(line)
        except Exception:
            continue
        if "fps" in d:
            print(tag, d["config"], round(d["fps"], 1), "fps", d["kernel_us"])
PY
