#!/bin/bash
# GPU box, one call: the round's evidence.  Outputs under gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
python tools/parity_report.py > gpurun_out/parity_report_r2.txt 2> gpurun_out/parity_report_r2.err
python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
python tools/bench_paths.py > gpurun_out/paths_r2.jsonl 2> gpurun_out/paths_r2.err
# launch list of the bench command (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/ncu_launches_r2.log 2>&1
# full captures: the surface kernel, then the K2 kernels
bash tools/ncu_full.sh r2_umma16 "umma16_grid" umma
for k in gp_eval_kernel gp_moments_kernel ts_gemm_nt; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/prof_r2_$k \
      python tools/bench_paths.py gp ts > gpurun_out/ncu_full_r2_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qei_kernel -s 2 -c 1 -f -o gpurun_out/prof_r2_qei_kernel \
    python tools/qei_profile.py > gpurun_out/ncu_full_r2_qei.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
