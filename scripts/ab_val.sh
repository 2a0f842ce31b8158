#!/bin/bash
# bench one config under several values of one environment variable: scripts/ab_val.sh <tag> <VAR> "<values>" "<cfgs>"
tag=$1; var=$2; vals=$3; cfgs=$4
for c in $cfgs; do
  for v in $vals; do
    if [ "$v" = "-" ]; then unset $var; else export $var=$v; fi
    timeout 300 python bench.py --config $c --single --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/${tag}_${c}_${var}_${v}.json 2> gpurun_out/${tag}_${c}_${v}.err || tail -3 gpurun_out/${tag}_${c}_${v}.err
    python - "gpurun_out/${tag}_${c}_${var}_${v}.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],4), d['gpu_launches']/d['steps'], d['roofline'].get('step_ms_by_call'))
except Exception as e: print(sys.argv[1],'ERR',e)
P
  done
done
unset $var
