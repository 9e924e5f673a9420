#!/bin/bash
# Feasibility: the existing feeder-specialised on-chip kernel on the 123-bus feeder (1 warp per SM)
mkdir -p gpurun_out
T=r02b
( time timeout 900 python bench.py --workload c4 --fbs-kernel 5 --steps 10 --no-cpu > gpurun_out/${T}_c4_jit.json 2> gpurun_out/${T}_c4_jit.err ) 2>&1 | grep real; tail -3 gpurun_out/${T}_c4_jit.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_c4_jit.json')); r=d['roofline']
print('value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],r['kernel'],r['bound'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gp_fbs_jit -c 1 -s 3 -o gpurun_out/${T}_jit_c4 python bench.py --workload c4 --fbs-kernel 5 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu1.log 2>&1; tail -2 gpurun_out/${T}_ncu1.log
