#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for mode in ${MODES:-pipelined collective pipelined collective}; do
  BOGP_BENCH_DEBUG=1 BOGP_EXCHANGE_MODE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/sq.json 2> gpurun_out/sq.err
  echo "== $mode"; grep "bench.py\[rank" gpurun_out/sq.err
  python - <<'PY'
import json
d = json.loads(open('gpurun_out/sq.json').read().strip().splitlines()[-1])
print('   step %.4f ms' % d['ms_per_step'], 'e2e %.4f ms' % d['e2e']['ms_per_step'])
PY
done
