#!/bin/bash
# experiments + evidence: update timings (phases, merged weight-gradient launch), generic kernel after the batched solve, long training runs
mkdir -p gpurun_out
GRX_PPO_TIMING=1 timeout 200 python tools/prof_update.py 4096 64 3 2>&1 | tail -4 | tee gpurun_out/r2o_time_update.log
echo "GRX_DW_MERGE=1" | tee -a gpurun_out/r2o_time_update.log
GRX_DW_MERGE=1 timeout 200 python tools/prof_update.py 4096 64 3 2>&1 | grep "graph replay" | tee -a gpurun_out/r2o_time_update.log
GRX_ENV_GENERIC=1 timeout 300 python tools/time_env.py 4096 60 | tee gpurun_out/r2o_time_envg.log
timeout 300 python -m pytest tests/test_physg_gpu.py tests/test_env_gpu.py -q -k "generic or full or physg or oracle" 2>&1 | tail -3 | tee gpurun_out/r2o_pytest_generic.log
timeout 900 python tools/train_log.py --robot GR1T1 --mesh heightfield --envs 4096 --iters 1500 --every 25 --out gpurun_out/r2o_train_hf_4096x1500.jsonl 2>&1 | tail -3
timeout 900 python tools/train_log.py --robot GR1T1 --mesh trimesh --envs 4096 --iters 600 --every 20 --out gpurun_out/r2o_train_trimesh_4096x600.jsonl 2>&1 | tail -3
timeout 900 python tools/train_log.py --robot GR1T1 --mesh plane --envs 2048 --iters 300 --every 10 --full-body --out gpurun_out/r2o_train_fullbody_plane_2048x300.jsonl 2>&1 | tail -3
