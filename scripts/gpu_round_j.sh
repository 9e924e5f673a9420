#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fbs_gpu.py -x -q -m gpu > gpurun_out/j_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/j_tests.log
for w in c2 fbs34; do
timeout 600 python bench.py --workload $w --steps 20 --no-cpu > gpurun_out/j_$w.json 2> gpurun_out/j_$w.err || tail gpurun_out/j_$w.err
python -c "
import json;d=json.load(open('gpurun_out/j_$w.json'));print('$w',d['value'],d['ms_per_step'],d['roofline']['kernel_ms'],'e2e',d['e2e'])"
done
