#!/bin/bash
# denoiser samples per launch sequence (dcb_max_batch) vs throughput on the bench workload
for mb in "$@"; do
  python bench.py --no-cpu --steps 3 --warmup 3 --max-batch $mb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('max_batch', $mb, 'evals/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'gemm TF/s', round(d['roofline']['achieved']), 'share', round(d['roofline']['kernel_share_of_step'],3), 'launches', d['gpu_launches'])"
done
