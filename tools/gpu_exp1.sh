#!/bin/bash
# experiments: (1) ncu --set full of the generic env kernel (lower-limb model on the generic path + full-body), (2) GEMM debug floors of the update
mkdir -p gpurun_out
GRX_ENV_GENERIC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:envg_step_kernel -s 25 -c 1 -f -o gpurun_out/r2k_envg python tools/time_env.py 1036 30 > gpurun_out/r2k_ncu_envg.log 2>&1
python tools/ncu_report.py gpurun_out/r2k_envg.ncu-rep > gpurun_out/r2k_envg_ncu_summary.txt; head -60 gpurun_out/r2k_envg_ncu_summary.txt
for d in 0 1 2 3; do echo "GRX_TC_DEBUG=$d"; GRX_TC_DEBUG=$d timeout 200 python tools/prof_update.py 4096 64 3 2>&1 | grep "graph replay"; done | tee gpurun_out/r2k_tc_debug.log
for t in 1x128 1x256 2x128 2x256; do echo "GRX_TC_TILE=$t"; GRX_TC_TILE=$t timeout 200 python tools/prof_update.py 4096 64 3 2>&1 | grep "graph replay"; done | tee -a gpurun_out/r2k_tc_debug.log
