#!/bin/bash
# A/B of the weight-stationary kernel against the slab kernel on one box: scripts/ws_ab.sh <tag>
tag=${1:-ws}
timeout 120 python scripts/ws_debug.py 1 > gpurun_out/${tag}_ws_debug.txt 2>&1; echo "ws_debug exit $?" >> gpurun_out/${tag}_ws_debug.txt
tail -3 gpurun_out/${tag}_ws_debug.txt
B="python bench.py --single --no-cpu --no-e2e --steps 20 --warmup 5"
NPML_WS=0 timeout 200 $B --config cfg2 > gpurun_out/${tag}_cfg2_ws0.json 2> gpurun_out/${tag}_err.log
NPML_WS=1 timeout 200 $B --config cfg2 > gpurun_out/${tag}_cfg2_ws1.json 2>> gpurun_out/${tag}_err.log
NPML_WS=1 NPML_WS_NT=128 timeout 200 $B --config cfg2 > gpurun_out/${tag}_cfg2_ws1_nt128.json 2>> gpurun_out/${tag}_err.log
NPML_WS=0 timeout 200 $B --config cfg4 > gpurun_out/${tag}_cfg4_ws0.json 2>> gpurun_out/${tag}_err.log
NPML_WS=1 timeout 200 $B --config cfg4 > gpurun_out/${tag}_cfg4_ws1.json 2>> gpurun_out/${tag}_err.log
python -  <<'P'
import json,glob,sys
for f in sorted(glob.glob('gpurun_out/%s_cfg*_ws*.json' % sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/*_ws*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']), d['ms_per_step'], d['roofline'].get('kernel_ms'))
    except Exception as e: print(f, 'ERR', e)
P
