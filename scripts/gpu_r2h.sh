#!/bin/bash
# NR3 phase-lane kernel: L2 prefetch distance sweep on C3 and C4-NR3; NR3 parity tests
mkdir -p gpurun_out
T=r02h
timeout 900 python -m pytest tests/test_nr3_gpu.py tests/test_big_feeders.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
for w in c3 c4nr3; do for pf in 0 1 2 3 4; do
  timeout 600 python bench.py --workload $w --prefetch $pf --steps 20 --no-cpu > gpurun_out/${T}_${w}_pf$pf.json 2> gpurun_out/${T}_${w}_pf$pf.err || tail -3 gpurun_out/${T}_${w}_pf$pf.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_c*_pf*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),r['kernel'],r['bound'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4))
    except Exception as e: print(f,'ERR',e)
PY
