#!/bin/bash
mkdir -p gpurun_out
for tl in 4 4; do
  BOGP_U16_TURN_LEFT=$tl python bench.py --steps 30 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/tl.json 2> gpurun_out/tl.err
  python - $tl <<'PY'
import json, sys
d = json.loads(open('gpurun_out/tl.json').read().strip().splitlines()[-1])
print('turn_left', sys.argv[1], '%.4g' % d['value'], 'step %.4f ms' % d['ms_per_step'], 'kernel %.4f' % d['roofline']['kernel_ms'], 'e2e %.4f ms' % d['e2e']['ms_per_step'], 'h2d', d['e2e']['h2d_bytes_per_step'], 'uncached %.4f' % d['e2e']['uncached']['ms_per_step'], d['e2e']['uncached']['h2d_bytes_per_step'])
PY
done
