#!/bin/bash
# device time per kernel of one config (ncu launch list, cold-cache): scripts/ncu_times.sh <tag> <cfg>
tag=$1; c=$2
cmd="python bench.py --config $c --single --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph --quick"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches_${c}.csv $cmd > gpurun_out/${tag}_ncu_${c}.log 2>&1
python - gpurun_out/${tag}_launches_${c}.csv <<'P'
import csv,sys,collections
d=collections.OrderedDict()
for r in csv.reader(open(sys.argv[1])):
    if len(r)>14 and r[0].isdigit():
        k=r[4].split('(')[0].replace('npml::<unnamed>::','').replace('void ','')[:70]
        d.setdefault(k,{}).setdefault(r[12],[]).append(float(r[14].replace(',','')))
for k,v in d.items():
    t=v.get('gpu__time_duration.sum',[0]); i=v.get('smsp__inst_executed.sum',[0])
    print('%-72s n=%3d  %8.1f us  %10.0f inst'%(k,len(t),sum(t)/len(t)/ (1000 if sum(t)/len(t)>1000 else 1),sum(i)/len(i)))
P
