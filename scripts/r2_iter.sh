#!/bin/bash
# one GPU iteration: the tests named in $2 (pytest -k expression), then a short bench of the configs in $3
tag=$1; kexpr=$2; cfgs=${3:-""}
timeout 900 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for c in $cfgs; do
  timeout 300 python bench.py --config $c --single --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/${tag}_${c}.json 2> gpurun_out/${tag}_${c}.err || tail -3 gpurun_out/${tag}_${c}.err
  python - "$tag" "$c" <<'P'
import json,sys
try:
    d=json.loads(open('gpurun_out/%s_%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d['value']), round(d['ms_per_step'],4), d['gpu_launches']/d['steps'], d['roofline'].get('kernel_ms'), d['roofline'].get('step_ms_by_call'))
except Exception as e: print(sys.argv[2],'ERR',e)
P
done
