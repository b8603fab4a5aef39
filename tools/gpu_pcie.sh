#!/bin/bash
mkdir -p gpurun_out
{
python tools/pcie_probe.py
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe.py 2>/dev/null
done
} > gpurun_out/pcie_probe_r2.log 2>&1
cat gpurun_out/pcie_probe_r2.log | grep "^N="
BOGP_NO_ZEROCOPY=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu --no-bo --no-paths > gpurun_out/scale_r2_n8_nozc.json 2> gpurun_out/scale_r2_n8_nozc.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/scale_r2_n8_nozc.json').read().strip().splitlines()[-1])
print('N=8 BOGP_NO_ZEROCOPY=1: step %.4f ms' % d['ms_per_step'], 'e2e %.4f ms' % d['e2e']['ms_per_step'])
PY
