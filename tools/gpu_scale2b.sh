#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for rep in 1 2; do
for mode in pipelined collective; do
for ts in 1 0; do
  BOGP_UMMA_TABSTREAM=$ts BOGP_EXCHANGE_MODE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/sq.json 2> gpurun_out/sq.err
  python - $mode $ts <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/sq.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'tabstream', sys.argv[2], d['n_gpus'], '%.4g' % d['value'], 'step %.4f ms' % d['ms_per_step'], 'kernel %.4f' % d['roofline']['kernel_ms'], 'e2e %.4f ms' % d['e2e']['ms_per_step'])
except Exception as e:
    print('ERR', e, open('gpurun_out/sq.err').read()[-500:])
PY
done; done; done
