#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nr3_gpu.py -x -q -m gpu > gpurun_out/n_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/n_tests.log
for w in c4nr3 nr3_123_big c3; do for kk in 1 2; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --nr3-kernel $kk > gpurun_out/n_${w}_$kk.json 2> gpurun_out/n_${w}_$kk.err || tail -3 gpurun_out/n_${w}_$kk.err
  python - <<PY
import json;d=json.load(open('gpurun_out/n_${w}_$kk.json'));print('$w kernel=$kk','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4))
PY
done; done
