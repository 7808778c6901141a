#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ppo_gpu.py tests/test_gemm_gpu.py -q -x 2>&1 | tail -3 | tee gpurun_out/r2s_pytest.log
timeout 200 python tools/prof_update.py 4096 64 3 2>&1 | grep "graph replay" | tee gpurun_out/r2s_time_update.log
timeout 900 python tools/train_log.py --robot GR1T1 --mesh plane --envs 2048 --iters 300 --every 10 --full-body --no-self-collision --out gpurun_out/r2s_train_fullbody_noselfcoll_plane_2048x300.jsonl 2>&1 | tail -2
