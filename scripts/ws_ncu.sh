#!/bin/bash
# ncu --set full capture (with source counters) of the weight-stationary kernel on cfg2: scripts/ws_ncu.sh <tag> [NT]
tag=${1:-ws}; nt=${2:-128}
export NPML_WS=1 NPML_WS_NT=$nt
cmd="python bench.py --config cfg2 --single --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph --quick"
$cmd > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_ws" -s 6 -c 2 -o gpurun_out/${tag}_ws -f $cmd > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
ls -la gpurun_out/${tag}_*
