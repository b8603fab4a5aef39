#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bartlett_gpu.py -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_ab_tab2.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/tab_probe.csv python tools/tab_probe.py 6 > gpurun_out/tab_probe.log 2>&1
python tools/summarize_launches.py gpurun_out/tab_probe.csv 2>/dev/null | grep umma16
