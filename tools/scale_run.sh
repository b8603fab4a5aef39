#!/bin/bash
# GPU box with N GPUs: bench.py at 1..N GPUs, weak and strong scaling, one summary line each.
#   gpurun --gpus 8 -- bash tools/scale_run.sh 8 tag
N=${1:-2}; TAG=${2:-scale}
mkdir -p gpurun_out
for mode in weak strong; do
  for n in 1 2 4 8; do
    [ $n -gt $N ] && continue
    out=gpurun_out/${TAG}_${mode}_n${n}.json
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --scaling $mode --no-paths --no-bo --no-cpu --steps 30 --warmup 5 > $out 2> gpurun_out/${TAG}_${mode}_n${n}.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
        bench.py --gpus $n --scaling $mode --no-paths --no-bo --no-cpu --steps 30 --warmup 5 > $out 2> gpurun_out/${TAG}_${mode}_n${n}.err
    fi
    python tools/print_bench.py $out "$mode n=$n" || tail -5 gpurun_out/${TAG}_${mode}_n${n}.err
  done
done
