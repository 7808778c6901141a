#!/bin/bash
# ncu launch list of one graph-replayed PPO update + env steps (per-kernel durations, cold cache).  Usage: tools/gpu_prof.sh <tag>
TAG=${1:-p}
mkdir -p gpurun_out
timeout 120 python tools/prof_update.py 4096 64 2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/${TAG}_upd_launches.csv python tools/prof_update.py 4096 64 2 > gpurun_out/${TAG}_ncu_upd.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_upd_launches.csv gpurun_out/${TAG}_upd_summary.csv | cut -c1-150 | head -30
