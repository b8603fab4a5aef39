#!/bin/bash
mkdir -p gpurun_out
for pre in 0 0 2 2; do
  BOGP_BENCH_PRESTEPS=$pre BOGP_BENCH_DEBUG=1 python bench.py --steps 20 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/sq.json 2> gpurun_out/sq.err
  echo "== presteps $pre"; grep "bench.py\[rank" gpurun_out/sq.err
  python - <<'PY'
import json
d = json.loads(open('gpurun_out/sq.json').read().strip().splitlines()[-1])
print('   step %.4f ms' % d['ms_per_step'], 'e2e %.4f ms' % d['e2e']['ms_per_step'])
PY
done
