set -x
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do
  timeout 200 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r60_${c}_bf16_bench.json 2> gpurun_out/r60_${c}_err.log
done
timeout 200 python bench.py --config cfg2 --mode tf32 --steps 20 --warmup 5 --no-cpu > gpurun_out/r60_cfg2_tf32_bench.json 2>> gpurun_out/r60_err.log
timeout 200 python bench.py --config cfg2 --mode fp32 --steps 5 --warmup 3 --no-cpu > gpurun_out/r60_cfg2_fp32_bench.json 2>> gpurun_out/r60_err.log
# launch lists (after the plain runs above exited 0)
for c in cfg2 cfg4 cfg5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r60_launches_$c.csv python bench.py --config $c --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/r60_ncu_$c.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_slab|wgrad_reduce" -s 9 -c 4 -o gpurun_out/r60_cfg2 -f python bench.py --steps 3 --warmup 3 --no-cpu --no-graph > gpurun_out/r60_ncu_full.log 2>&1
ls -la gpurun_out/r60*
