#!/bin/bash
# ncu campaign of one round: for every config a launch list (device time per launch) and an `--set full` capture of its
# dominant kernels, each only after the same command has exited 0 without ncu.  Reports land in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r2 "cfg2 cfg4"'
tag=${1:-r2}
cfgs=${2:-"cfg1 cfg2 cfg3 cfg4 cfg5"}
for c in $cfgs; do
  cmd="python bench.py --config $c --single --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph --quick"
  $cmd > gpurun_out/${tag}_${c}_plain.log 2>&1 || { echo "$c: plain run failed"; tail -3 gpurun_out/${tag}_${c}_plain.log; continue; }
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_${c}.csv $cmd > gpurun_out/${tag}_ncu_${c}.log 2>&1
  case $c in
    cfg1) pat="direct|thin|reduce"; cnt=6;;
    cfg2) pat="tc_slab|wgrad_reduce"; cnt=4;;
    cfg3) pat="tc_slab|tc_gather|tc_wgrad|wgrad_reduce"; cnt=5;;
    cfg4) pat="tc_slab|wgrad_reduce|col_pass"; cnt=20;;
    cfg5) pat="tc_slab|tc_wgrad|wgrad_reduce|col_pass"; cnt=6;;
  esac
  # skip the launches of warm-up (3 steps) so that the captured ones belong to a steady-state step
  nwarm=$(python scripts/ncu_summary.py count gpurun_out/${tag}_launches_${c}.csv "$pat" 2>/dev/null || echo 0)
  skip=$(( nwarm * 5 / 8 ))
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s $skip -c $cnt -o gpurun_out/${tag}_${c} -f $cmd > gpurun_out/${tag}_ncu_full_${c}.log 2>&1
  # the reports are large (--import-source): bring back the raw metric table of every config, the report itself only
  # for the headline config (gpurun merges at most 64 MiB)
  if [ -f gpurun_out/${tag}_${c}.ncu-rep ]; then
    ncu -i gpurun_out/${tag}_${c}.ncu-rep --page raw --csv > gpurun_out/${tag}_${c}_raw.csv 2>/dev/null
    if [ "$c" != "cfg2" ]; then rm -f gpurun_out/${tag}_${c}.ncu-rep; fi
  fi
  ls -la gpurun_out/${tag}_${c}* 2>/dev/null
done
