timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | tail -4
for ov in 1 0; do
GRX_COMM_OVERLAP=$ov timeout 600 python bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('overlap=$ov', 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'coll', round(d['config']['collection_ms'],2), 'learn', round(d['config']['learn_ms'],2), 'comm_error', d['comm_error'], 'identical', d['replicas_identical'], 'launches', d['gpu_launches'])"
done
