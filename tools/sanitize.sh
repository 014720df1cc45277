#!/bin/bash
# compute-sanitizer over a handful of small frames of every kernel family (tools/sanitize_frame.py): memcheck, racecheck
# (shared memory hazards), synccheck (barriers / mbarriers), initcheck (uninitialised global reads). One gpurun call:
#   tools/sanitize.sh <tag>      -> gpurun_out/<tag>_sanitize_<tool>.log
tag=${1:-r00}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  SAN_SPLATS=${SAN_SPLATS:-20000} timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_frame.py > gpurun_out/${tag}_sanitize_${tool}.log 2>&1
  echo "== $tool: rc=$? $(grep -c 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/${tag}_sanitize_${tool}.log)"; grep "SUMMARY\|sanitize_frame ok" gpurun_out/${tag}_sanitize_${tool}.log | tail -3
done
