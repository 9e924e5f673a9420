#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nr3_gpu.py tests/test_big_feeders.py -x -q -m gpu > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/i_tests.log
for w in c4nr3 nr3_123_big c3; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/i_${w}.json 2> gpurun_out/i_${w}.err || tail -3 gpurun_out/i_${w}.err
  python - <<PY
import json;d=json.load(open('gpurun_out/i_${w}.json'));print('$w','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4))
PY
done
for w in fbs34 fbs37; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --fbs-kernel 5 > gpurun_out/i_${w}_5.json 2> gpurun_out/i_${w}_5.err || tail -3 gpurun_out/i_${w}_5.err
  python - <<PY
import json;d=json.load(open('gpurun_out/i_${w}_5.json'));print('$w kernel 5','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4))
PY
done
