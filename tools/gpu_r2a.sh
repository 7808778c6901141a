#!/bin/bash
# Round-2 call A (1 GPU): parity tests, training evidence (300 PPO iterations, plane + heightfield), bench of the three configs.
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/train_log.py --mesh heightfield --iters 300 --out gpurun_out/${TAG}_train_hf.jsonl > gpurun_out/${TAG}_train_hf.log 2>&1; tail -2 gpurun_out/${TAG}_train_hf.log
timeout 600 python tools/train_log.py --mesh plane --iters 300 --out gpurun_out/${TAG}_train_plane.jsonl > gpurun_out/${TAG}_train_plane.log 2>&1; tail -1 gpurun_out/${TAG}_train_plane.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err; tail -c 2500 gpurun_out/${TAG}_bench_c2.json
