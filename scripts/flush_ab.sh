#!/bin/bash
# A/B of the two L2 flush styles on the small configs: scripts/flush_ab.sh <tag>
tag=$1
for f in write write+read; do
  for c in cfg1 cfg3 cfg5; do
    BENCH_FLUSH=$f timeout 300 python bench.py --config $c --single --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/${tag}_${c}_${f}.json 2> gpurun_out/${tag}_${c}_${f}.err || tail -3 gpurun_out/${tag}_${c}_${f}.err
    python - "gpurun_out/${tag}_${c}_${f}.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],4), d['roofline'].get('kernel_ms'))
except Exception as e: print(sys.argv[1],'ERR',e)
P
  done
done
