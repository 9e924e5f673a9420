#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fbs_gpu.py -x -q -m gpu > gpurun_out/m_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/m_tests.log
for w in fbs37; do for ig in -1; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --fbs-kernel 5 --jit-ig $ig > gpurun_out/m_${w}_$ig.json 2> gpurun_out/m_${w}_$ig.err || tail -3 gpurun_out/m_${w}_$ig.err
  python - <<PY
import json;d=json.load(open('gpurun_out/m_${w}_$ig.json'));print('$w ig=$ig','value %.3e'%d['value'],'ms',round(d['ms_per_step'],4),'kernel_ms',round(d['roofline']['kernel_ms'],4),'e2e %.3e'%d['e2e']['value'])
PY
done; done
