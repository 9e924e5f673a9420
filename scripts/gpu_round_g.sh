#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nr3_gpu.py -x -q -m gpu > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/g_tests.log
for w in c3 c4nr3 nr3_37_big; do for pf in 0 2; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --prefetch $pf > gpurun_out/g_${w}_$pf.json 2> gpurun_out/g_${w}_$pf.err || tail -3 gpurun_out/g_${w}_$pf.err
  python - <<PY
import json;d=json.load(open('gpurun_out/g_${w}_$pf.json'));print('$w pf=$pf','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4))
PY
done; done
