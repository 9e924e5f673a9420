#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nr3_gpu.py tests/test_big_feeders.py tests/test_study_gpu.py -x -q -m gpu > gpurun_out/l_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/l_tests.log
for w in c3 c4nr3 nr3_37_big nr3_123_big; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu > gpurun_out/l_${w}.json 2> gpurun_out/l_${w}.err || tail -3 gpurun_out/l_${w}.err
  python - <<PY
import json;d=json.load(open('gpurun_out/l_${w}.json'));print('$w','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4), 'it', d['config']['mean_iterations'])
PY
done
