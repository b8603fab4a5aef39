#!/bin/bash
# Usage (on the GPU box, via gpurun): bash tools/gpu_round.sh <tag> [engine]
# GPU parity tests, smoke, bench (N=1), then the ncu launch list and one full capture of the dominant kernel.
tag=${1:-r1}; engine=${2:-auto}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/gputests_$tag.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --engine $engine > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$tag.json
timeout 300 python bench.py --engine $engine --steps 3 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/plain_$tag.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --engine $engine --steps 3 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/ncu_list_$tag.log 2>&1
echo "ncu list rc=$?"
