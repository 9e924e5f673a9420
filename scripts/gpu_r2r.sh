#!/bin/bash
# depth-first build of the specialised FBS kernel: parity tests, A/B on the 34- and 37-bus feeders and C2
mkdir -p gpurun_out
T=r02r
timeout 900 python -m pytest tests/test_fbs_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/${T}_tests.log
for w in fbs34 fbs37 c2; do for d in 0 1; do
  timeout 600 python bench.py --workload $w --jit-dfs $d --steps 20 --no-cpu > gpurun_out/${T}_${w}_dfs$d.json 2> gpurun_out/${T}_${w}_dfs$d.err || tail -3 gpurun_out/${T}_${w}_dfs$d.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*_dfs*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],4),r['kernel'],r['bound'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4),'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
PY
