#!/bin/bash
mkdir -p gpurun_out
for ts in 1 0 1 0; do
  BOGP_UMMA_TABSTREAM=$ts python bench.py --steps 30 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/tab_$ts.json 2> gpurun_out/tab_$ts.err
  python - $ts <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/tab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('tabstream', sys.argv[1], '%.4g' % d['value'], 'step %.4f ms' % d['ms_per_step'], 'kernel %.4f' % d['roofline']['kernel_ms'], 'e2e %.4f ms' % d['e2e']['ms_per_step'])
PY
done
