#!/bin/bash
mkdir -p gpurun_out
T=r02n
for w in c5; do for pf in 0 1 2 4; do
  timeout 600 python bench.py --workload $w --prefetch $pf --steps 10 --no-cpu > gpurun_out/${T}_${w}_pf$pf.json 2> gpurun_out/${T}_${w}_pf$pf.err || tail -3 gpurun_out/${T}_${w}_pf$pf.err
done; done
timeout 600 python bench.py --workload c4 --steps 20 --no-cpu > gpurun_out/${T}_c4_auto.json 2> gpurun_out/${T}_c4_auto.err
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_c*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),r['kernel'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4))
    except Exception as e: print(f,'ERR',e)
PY
