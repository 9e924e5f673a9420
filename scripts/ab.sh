#!/bin/bash
# A/B the two FBS kernels on the three FBS workloads (no CPU leg)
for w in c2 c4 c5; do for k in 0 2; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --fbs-kernel $k > gpurun_out/ab_${w}_$k.json 2> gpurun_out/ab_${w}_$k.err || tail -3 gpurun_out/ab_${w}_$k.err
  python -c "
import json;d=json.load(open('gpurun_out/ab_${w}_$k.json'));print('$w kernel=$k','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3))"
done; done
