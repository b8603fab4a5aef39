#!/bin/bash
# GPU box with >= 2 GPUs: peer-exchange tests, then the weak-scaling bench at N = 1 and N = 2 with the pipelined and the
# collective form of the final reduction.
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests/test_peer.py -m gpu -x -q > gpurun_out/peer_tests_n$N.log 2>&1; tail -3 gpurun_out/peer_tests_n$N.log
python bench.py --steps 30 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/sp_n1.json 2> gpurun_out/sp_n1.err
for mode in pipelined collective; do
  BOGP_EXCHANGE_MODE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/sp_${mode}_n$N.json 2> gpurun_out/sp_${mode}_n$N.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sp_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], '%.4g' % d['value'], 'step %.4f ms' % d['ms_per_step'], 'e2e %.4f ms' % d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
