#!/bin/bash
# GPU session: C4 / C5 streaming kernel profile
mkdir -p gpurun_out
for w in c4 c5; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/c_$w.json 2> gpurun_out/c_$w.err || tail -3 gpurun_out/c_$w.err
  python - <<PY
import json;d=json.load(open('gpurun_out/c_$w.json'));print('$w','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4), d['roofline']['bound'], d['config']['fbs_kernel'])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbs_solve_kernel -c 1 -s 4 -o gpurun_out/c_stream_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/c_ncu1.log 2>&1; tail -2 gpurun_out/c_ncu1.log
