#!/bin/bash
# one GPU: Bartlett / peer / front-end tests with the streamed tables, then same-box A/B of the table paths
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bartlett_gpu.py tests/test_peer.py tests/test_frontend.py -m gpu -x -q > gpurun_out/tab_tests.log 2>&1; tail -3 gpurun_out/tab_tests.log
for ts in 1 0 1 0; do
  BOGP_UMMA_TABSTREAM=$ts python bench.py --steps 30 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/tab_$ts.json 2> gpurun_out/tab_$ts.err
  python - $ts <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/tab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('tabstream', sys.argv[1], '%.4g' % d['value'], 'step %.4f ms' % d['ms_per_step'], 'kernel %.4f' % d['roofline']['kernel_ms'], 'e2e %.4f ms' % d['e2e']['ms_per_step'])
PY
done
