#!/bin/bash
for v in 0 1 2 3; do echo "GRX_LAYER_PIPE=$v"; GRX_LAYER_PIPE=$v timeout 200 python tools/prof_update.py 4096 64 3 2>&1 | grep "graph replay"; done | tee gpurun_out/r2r_pipe_variants.log
