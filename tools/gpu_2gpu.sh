#!/bin/bash
# 2-GPU round: the multi-GPU parity test + bench at N = 2 (launched like the driver does)
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -q -s > gpurun_out/${TAG}_multigpu_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_multigpu_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err; tail -c 1200 gpurun_out/${TAG}_bench_2gpu.json; tail -3 gpurun_out/${TAG}_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --config 5 > gpurun_out/${TAG}_bench_2gpu_c5.json 2> gpurun_out/${TAG}_bench_2gpu_c5.err; tail -c 600 gpurun_out/${TAG}_bench_2gpu_c5.json
