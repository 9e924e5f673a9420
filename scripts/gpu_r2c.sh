#!/bin/bash
# Depth-first order of the streaming kernel: parity tests, then A/B on C4 and C5
mkdir -p gpurun_out
T=r02c
timeout 900 python -m pytest tests/test_fbs_gpu.py tests/test_big_feeders.py tests/test_extras_gpu.py tests/test_study_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
for w in c4 c5; do for o in 0 1; do
  timeout 600 python bench.py --workload $w --fbs-order $o --steps 10 --no-cpu > gpurun_out/${T}_${w}_o$o.json 2> gpurun_out/${T}_${w}_o$o.err || tail -3 gpurun_out/${T}_${w}_o$o.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_c*_o*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],r['kernel'],r['bound'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4))
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbs_solve_kernel -c 1 -s 3 -o gpurun_out/${T}_stream_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu1.log 2>&1; tail -2 gpurun_out/${T}_ncu1.log
