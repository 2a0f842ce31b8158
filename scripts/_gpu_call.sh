timeout 120 python scripts/ws_debug.py > gpurun_out/r80_ws_debug.txt 2>&1; echo "ws_debug exit $?" >> gpurun_out/r80_ws_debug.txt
tail -12 gpurun_out/r80_ws_debug.txt | cut -c1-200
if grep -q "all cases agree" gpurun_out/r80_ws_debug.txt; then
  NPML_WS=1 timeout 120 python bench.py --config cfg2 --steps 30 --warmup 5 --no-cpu > gpurun_out/r80_cfg2_ws1.json 2>> gpurun_out/r80_err.log
  NPML_WS=1 NPML_WS_NT=128 timeout 120 python bench.py --config cfg2 --steps 30 --warmup 5 --no-cpu > gpurun_out/r80_cfg2_ws1_nt128.json 2>> gpurun_out/r80_err.log
  NPML_WS=1 NPML_WS_NT=64 timeout 120 python bench.py --config cfg2 --steps 30 --warmup 5 --no-cpu > gpurun_out/r80_cfg2_ws1_nt64.json 2>> gpurun_out/r80_err.log
  NPML_WS=1 timeout 120 python bench.py --config cfg4 --steps 30 --warmup 5 --no-cpu > gpurun_out/r80_cfg4_ws1.json 2>> gpurun_out/r80_err.log
  python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r80_cfg*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'ERR', e)
PY
  tail -5 gpurun_out/r80_err.log
fi
