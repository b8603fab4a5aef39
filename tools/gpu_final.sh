#!/bin/bash
# GPU box, one call: the final evidence of the round with the library as committed.
mkdir -p gpurun_out
python tools/parity_report.py > gpurun_out/parity_report_r2.txt 2> gpurun_out/parity_report_r2.err
python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
python tools/bench_paths.py > gpurun_out/paths_r2.jsonl 2> gpurun_out/paths_r2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/ncu_launches_r2.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:umma16_grid -s 3 -c 2 -f -o gpurun_out/prof_r2_umma16 \
    python bench.py --engine umma --steps 2 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/ncu_full_r2_umma16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_fit_coop -c 1 -f -o gpurun_out/prof_r2_gp_fit_coop \
    python tools/fit_profile.py 256 > gpurun_out/ncu_full_r2_fit.log 2>&1
BOGP_GP_DMMA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_r2_gp_eval_kernel \
    python tools/bench_paths.py gp > gpurun_out/ncu_full_r2_gp_eval.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8; tail -3 gpurun_out/parity_report_r2.txt | cut -c1-200
