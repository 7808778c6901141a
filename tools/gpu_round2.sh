#!/bin/bash
# One gpurun call of round 2: the whole GPU test suite, bench lines of configs 2 / 3 / 5 / 6, env timings (fused lower-limb kernel, generic kernels).
TAG=${1:-r}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for c in 2 3 5 6; do
  timeout 600 python bench.py --steps 10 --warmup 3 --config $c > gpurun_out/${TAG}_bench_c$c.json 2> gpurun_out/${TAG}_bench_c$c.err; tail -c 600 gpurun_out/${TAG}_bench_c$c.json
done
timeout 300 python tools/time_env.py 4096 200 | tee gpurun_out/${TAG}_time_env.log
GRX_ENV_GENERIC=1 timeout 300 python tools/time_env.py 4096 100 | tee -a gpurun_out/${TAG}_time_env.log
timeout 300 python tools/prof_update.py 4096 64 3 | tee gpurun_out/${TAG}_time_update.log
ls -la gpurun_out | tail -8
