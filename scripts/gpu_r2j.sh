#!/bin/bash
# NR3 phase-lane kernel with the cp.async ring: parity tests, A/B on C3 and C4-NR3
mkdir -p gpurun_out
T=r02j
timeout 900 python -m pytest tests/test_nr3_gpu.py tests/test_big_feeders.py tests/test_study_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
for w in c3 c4nr3; do for r in 0 1; do
  timeout 600 python bench.py --workload $w --nr3-ring $r --steps 20 --no-cpu > gpurun_out/${T}_${w}_ring$r.json 2> gpurun_out/${T}_${w}_ring$r.err || tail -3 gpurun_out/${T}_${w}_ring$r.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_c*_ring*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),r['kernel'],r['bound'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4))
    except Exception as e: print(f,'ERR',e)
PY
