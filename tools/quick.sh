#!/bin/bash
# GPU box: the Bartlett parity tests, then the config-2 bench line twice (summary only).  Usage: bash tools/quick.sh [tag]
timeout 400 python -m pytest tests/test_bartlett_gpu.py -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do
  timeout 200 python bench.py --no-paths --no-bo --no-cpu --steps 30 --warmup 5 2>gpurun_out/quick.err | python tools/print_bench.py - "$1" || tail -5 gpurun_out/quick.err
done
