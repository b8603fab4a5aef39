#!/bin/bash
# GPU box, one call: the final evidence of the round with the library as committed.  Order matters: the ncu capture of the
# surface kernel comes first, its DRAM bytes go into profiles/umma_traffic.json together with the library digest, and the
# bench line that follows reads them (and refuses them under any other build).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:umma16_grid -s 3 -c 2 -f -o gpurun_out/prof_r2_umma16 \
    python bench.py --engine umma --steps 2 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/ncu_full_r2_umma16.log 2>&1
python tools/make_traffic.py gpurun_out/prof_r2_umma16.ncu-rep 16 > gpurun_out/make_traffic.log 2>&1 && cp profiles/umma_traffic.json gpurun_out/umma_traffic.json
python tools/ncu_summary.py gpurun_out/prof_r2_umma16.ncu-rep > gpurun_out/ncu_r2_umma16.md 2> gpurun_out/ncu_summary.err
python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
python tools/parity_report.py > gpurun_out/parity_report_r2.txt 2> gpurun_out/parity_report_r2.err
python tools/bench_paths.py > gpurun_out/paths_r2.jsonl 2> gpurun_out/paths_r2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-bo --no-paths > gpurun_out/ncu_launches_r2.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2.csv > gpurun_out/launches_r2.md 2>/dev/null
ls -la gpurun_out/*.ncu-rep | tail -3; tail -3 gpurun_out/parity_report_r2.txt | cut -c1-200; cut -c1-400 gpurun_out/bench_r2f.json; cat gpurun_out/make_traffic.log | cut -c1-300
