#!/bin/bash
# e2e (host buffers) throughput of cfg2 as a function of the host pipeline's chunk count: scripts/e2e_chunks.sh <tag>
tag=$1
for ch in 4 8 16 32; do
  timeout 300 python bench.py --config cfg2 --single --no-cpu --no-dropin --steps 10 --warmup 3 --chunks $ch > gpurun_out/${tag}_chunks${ch}.json 2> gpurun_out/${tag}_chunks${ch}.err || tail -3 gpurun_out/${tag}_chunks${ch}.err
  python - "gpurun_out/${tag}_chunks${ch}.json" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value']), 'e2e f32', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'e2e bf16', round(d.get('e2e_bf16_host',{}).get('value',0)), d.get('e2e_bf16_host',{}).get('ms_per_step'))
P
done
