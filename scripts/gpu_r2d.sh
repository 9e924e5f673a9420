#!/bin/bash
# Tree-walk kernel: parity tests, then register-budget A/B on C4 and C5
mkdir -p gpurun_out
T=r02e
timeout 900 python -m pytest tests/test_fbs_gpu.py tests/test_big_feeders.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/${T}_tests.log
for w in c4 c5; do for b in 4 5 6 8; do
  timeout 600 python bench.py --workload $w --fbs-kernel 6 --walk-blocks $b --steps 10 --no-cpu > gpurun_out/${T}_${w}_b$b.json 2> gpurun_out/${T}_${w}_b$b.err || tail -3 gpurun_out/${T}_${w}_b$b.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_c*_b*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],r['kernel'],r['bound'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4))
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbs_walk_kernel -c 1 -s 3 -o gpurun_out/${T}_walk_c4 python bench.py --workload c4 --fbs-kernel 6 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu1.log 2>&1; tail -2 gpurun_out/${T}_ncu1.log
