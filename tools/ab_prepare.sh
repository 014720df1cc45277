#!/bin/bash
# A/B two builds in ONE gpurun call. In the build container:
#     tools/ab_prepare.sh <commit>          # checks <commit> out into ab_wt/ (git-ignored, but shipped to the GPU box) and builds it
#     gpurun -- 'tools/ab_run.sh 2 gut'     # bench_configs rows of this checkout, then of ab_wt/
#     tools/ab_prepare.sh --clean           # removes the worktree again
set -e
cd "$(dirname "$0")/.."
if [ "$1" == "--clean" ]; then git worktree remove --force ab_wt 2>/dev/null || true; git worktree prune; exit 0; fi
[ -n "$1" ] || { echo "usage: $0 <commit> | --clean"; exit 2; }
git worktree remove --force ab_wt 2>/dev/null || true
git worktree add -f ab_wt "$1" -q
# older checkouts hard-code the import path of the tool
sed -i "s#sys.path.insert(0, '/root/repo')#sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))#" ab_wt/tools/bench_configs.py
(cd ab_wt && python -m vk_gaussian_splatting_b200.build | tail -1)
ls -la ab_wt/vk_gaussian_splatting_b200/lib/
