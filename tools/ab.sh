#!/bin/bash
# Same-box A/B of two builds of libbogp_sm100.so (boxes of the pool differ by +-4 %, more than most kernel changes).
#
# Here (build container):   bash tools/ab.sh prepare          # builds HEAD -> ab/old.so and the working tree -> ab/new.so
# On the GPU box:           gpurun -- 'bash tools/ab.sh run'   # alternates old / new, prints step / kernel / e2e ms
# The .so files are swapped under an unchanged source tree, so bogp_b200.build sees a valid stamp and does not rebuild.
set -e
cd "$(dirname "$0")/.."
case "$1" in
  prepare)
    mkdir -p ab
    if git diff --quiet; then echo "working tree has no changes against HEAD: nothing to compare"; exit 1; fi
    git stash -q
    python -c "from bogp_b200 import build; build.build()" > /dev/null
    cp bogp_b200/libbogp_sm100.so ab/old.so
    git stash pop -q
    python -c "from bogp_b200 import build; build.build()" > /dev/null
    cp bogp_b200/libbogp_sm100.so ab/new.so
    echo "ab/old.so = HEAD, ab/new.so = working tree" ;;
  run)
    for v in old new old new; do
      cp ab/$v.so bogp_b200/libbogp_sm100.so
      python bench.py --no-paths --no-bo --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'step', round(d['ms_per_step'], 4), 'kernel', round(d['roofline']['kernel_ms'], 4), 'e2e', round(d['e2e']['ms_per_step'], 4))"
    done
    cp ab/new.so bogp_b200/libbogp_sm100.so ;;
  *) echo "usage: tools/ab.sh prepare | run"; exit 1 ;;
esac
