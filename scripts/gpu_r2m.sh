#!/bin/bash
# streaming FBS kernel: L1-prefetch mode (fbs_prefetch 100 + d) on C4 and C5
mkdir -p gpurun_out
T=r02m
for w in c4 c5; do for pf in 0 101 102 103; do
  timeout 600 python bench.py --workload $w --prefetch $pf --steps 10 --no-cpu > gpurun_out/${T}_${w}_pf$pf.json 2> gpurun_out/${T}_${w}_pf$pf.err || tail -3 gpurun_out/${T}_${w}_pf$pf.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_c*_pf*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),r['kernel'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4))
    except Exception as e: print(f,'ERR',e)
PY
