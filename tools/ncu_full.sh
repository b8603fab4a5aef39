#!/bin/bash
# Usage (GPU box): bash tools/ncu_full.sh <tag> <kernel-regex> [engine]   -- one `ncu --set full` capture of the dominant kernel
tag=$1; pat=$2; engine=${3:-auto}
mkdir -p gpurun_out
timeout 300 python bench.py --engine $engine --steps 2 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/plain_full_$tag.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 2 -f -o gpurun_out/prof_$tag \
    python bench.py --engine $engine --steps 2 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_$tag.log
