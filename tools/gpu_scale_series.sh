#!/bin/bash
# 8-GPU box: the weak-scaling series of the round (N = 1, 2, 4, 8, default settings) and the N = 8 run without the NUMA binding
mkdir -p gpurun_out
run() {  # N tag [env...]
  local N=$1 tag=$2; shift 2
  if [ "$N" = 1 ]; then
    env "$@" python bench.py --steps 20 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/scale_r2_$tag.json 2> gpurun_out/scale_r2_$tag.err
  else
    env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/scale_r2_$tag.json 2> gpurun_out/scale_r2_$tag.err
  fi
  python - $tag <<'PY'
import json, sys
try:
    d = json.loads(open(f'gpurun_out/scale_r2_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], d['n_gpus'], '%.4g' % d['value'], 'step %.4f ms' % d['ms_per_step'], 'e2e %.4g (%.4f ms)' % (d['e2e']['value'], d['e2e']['ms_per_step']), 'uncached %.4f' % d['e2e']['uncached']['ms_per_step'], 'numa', d['e2e'].get('numa_node'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
run 8 n8 BOGP_BENCH_DEBUG=0
run 8 n8_nonuma BOGP_BENCH_NUMA=0
run 4 n4 BOGP_BENCH_DEBUG=0
run 2 n2 BOGP_BENCH_DEBUG=0
run 1 n1 BOGP_BENCH_DEBUG=0
