#!/bin/bash
# fp32 exactness mode: parity subset, then cfg2 / cfg4 with and without the register-tiled CUDA-core kernels
tag=$1
timeout 900 python -m pytest tests -m gpu -x -q -k "fp32" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
for v in on off; do
  if [ $v = off ]; then export NPML_DISABLE_FP32_RT=1; else unset NPML_DISABLE_FP32_RT; fi
  for c in cfg2 cfg3 cfg5; do
    timeout 300 python bench.py --config $c --mode fp32 --single --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/${tag}_${c}_fp32_$v.json 2> gpurun_out/${tag}_${c}_$v.err || tail -3 gpurun_out/${tag}_${c}_$v.err
    python - "gpurun_out/${tag}_${c}_fp32_$v.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), d['roofline'].get('step_ms_by_call'))
except Exception as e: print(sys.argv[1],'ERR',e)
P
  done
done
