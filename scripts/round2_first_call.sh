#!/bin/bash
# First GPU call of the next round (everything here was prepared at the end of round 1 after the GPU budget was spent):
#   1. scripts/tma_store_probe.cu -- do TMA tensor stores clip a box that hangs over the tensor and un-swizzle by
#      absolute shared-memory address?  (the two assumptions of csrc/tc_ws2_conv.cu)
#   2. scripts/ws_debug.py 2      -- the experimental NPML_WS=2 kernel against the slab kernel, per tap when wrong
#   3. A/B of cfg2 / cfg4: default, NPML_WS=1 (validated), NPML_WS=2
# Build the probes on the CPU box first (they travel in build/):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tma_store_probe scripts/tma_store_probe.cu -lcuda -L/usr/local/cuda/lib64/stubs
# Usage: gpurun --timeout 300 -- 'bash scripts/round2_first_call.sh r2a'
tag=${1:-r2a}
o=gpurun_out/${tag}
timeout 30 build/tma_store_probe > ${o}_tma_store_probe.txt 2>&1; echo "exit $?" >> ${o}_tma_store_probe.txt
cat ${o}_tma_store_probe.txt | cut -c1-200
timeout 120 python scripts/ws_debug.py 2 > ${o}_ws2_debug.txt 2>&1; echo "exit $?" >> ${o}_ws2_debug.txt
tail -30 ${o}_ws2_debug.txt | cut -c1-300
for cfg in cfg2 cfg4; do
  for ws in 0 1 2; do
    if [ $ws = 2 ] && ! grep -q "all cases agree" ${o}_ws2_debug.txt; then continue; fi
    NPML_WS=$ws timeout 120 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu > ${o}_${cfg}_ws${ws}.json 2>> ${o}_err.log
  done
done
python - "$o" <<'PY'
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + '_cfg*_ws*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'], 4), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e:
        print(f, 'ERR', e)
PY
