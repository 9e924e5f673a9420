#!/bin/bash
mkdir -p gpurun_out
for w in c4 c5 c2; do for pf in 0 1 2 4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --fbs-kernel 3 --prefetch $pf > gpurun_out/d_${w}_$pf.json 2> gpurun_out/d_${w}_$pf.err || tail -3 gpurun_out/d_${w}_$pf.err
  python - <<PY
import json;d=json.load(open('gpurun_out/d_${w}_$pf.json'));print('$w pf=$pf','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4))
PY
done; done
