#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list and a full capture of one frame.
# usage: tools/gpu_round.sh <tag>
tag=${1:-r01x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip 20 --launch-count 10 -f -o gpurun_out/${tag}_full \
    python tools/prof_frame.py 4 cfg2 4 > gpurun_out/${tag}_ncu_full.log 2>&1
python tools/bench_configs.py gut gutx surf 2 3 5 sort > gpurun_out/${tag}_configs.jsonl 2>&1
python tools/benchmark_3dgs.py > gpurun_out/${tag}_benchmark_3dgs.log 2>&1
# the C++ host drivers over the C ABI (include/vkgs_b200.hpp)
L=$PWD/vk_gaussian_splatting_b200/lib
g++ -std=c++17 -Iinclude examples/render_host.cpp -L$L -lvkgs_b200 -Wl,-rpath,$L -o gpurun_out/render_host \
  && gpurun_out/render_host --synth 1000000 --ftb --frames 20 --ppm gpurun_out/${tag}_host.ppm > gpurun_out/${tag}_render_host.log 2>&1
g++ -std=c++17 -pthread -Iinclude examples/farm_host.cpp -L$L -lvkgs_b200 -Wl,-rpath,$L -o gpurun_out/farm_host \
  && gpurun_out/farm_host --gpus 1 --frames 200 > gpurun_out/${tag}_farm_host.log 2>&1
rm -f gpurun_out/render_host gpurun_out/farm_host gpurun_out/${tag}_host.ppm
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_smoke.log; cat gpurun_out/${tag}_farm_host.log; cut -c1-400 gpurun_out/${tag}_bench.json
