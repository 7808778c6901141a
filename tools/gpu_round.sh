#!/bin/bash
# One gpurun call: parity tests, bench, ncu launch list of the bench command, ncu --set full captures of the env kernel and of one
# minibatch's kernels, text summaries for profiles/.  Usage (on the GPU box, via gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json
timeout 300 python tools/time_env.py 4096 200 | tee gpurun_out/${TAG}_time_env.log
timeout 300 python tools/prof_update.py 4096 64 3 | tee gpurun_out/${TAG}_time_update.log
# one PPO iteration = 64 x 8 + 6 + 200 x 10 kernels; skip the first learn(1) + one warm-up iteration
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 2400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_launches_summary.csv | head -16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_step_kernel -s 30 -c 1 -f -o gpurun_out/${TAG}_env python tools/time_env.py 4096 40 > gpurun_out/${TAG}_ncu_env.log 2>&1
# the rollout phase launches 64 x 3 + 3 grouped GEMMs before the first minibatch of the update
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32|heads|apply_kernel|gather' -s 195 -c 11 -f -o gpurun_out/${TAG}_upd python tools/prof_update.py 4096 64 2 > gpurun_out/${TAG}_ncu_upd.log 2>&1
python tools/ncu_report.py gpurun_out/${TAG}_env.ncu-rep > gpurun_out/${TAG}_env_ncu_summary.txt
python tools/ncu_report.py gpurun_out/${TAG}_upd.ncu-rep > gpurun_out/${TAG}_update_kernels_ncu_summary.txt
ls -la gpurun_out | tail -15
