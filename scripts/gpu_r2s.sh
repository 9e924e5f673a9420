#!/bin/bash
# host path of feeders whose device kernel is the depth-first build: parity + e2e on the 34- and 37-bus feeders
mkdir -p gpurun_out
T=r02s
timeout 600 python -m pytest tests/test_fbs_gpu.py -x -q -m gpu -k "depth_first or one_launch or i_rows" > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/${T}_tests.log
for w in fbs34 fbs37; do
  timeout 600 python bench.py --workload $w --steps 20 --no-cpu > gpurun_out/${T}_${w}.json 2> gpurun_out/${T}_${w}.err || tail -3 gpurun_out/${T}_${w}.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_fbs*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],4),r['kernel'],r['bound'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4),'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
PY
