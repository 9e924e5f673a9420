#!/bin/bash
# GPU session: FBS parity tests incl. the feeder-specialised kernel, A/B on C2, latency probe, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fbs_gpu.py -x -q -m gpu > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/b_tests.log
for k in 0; do
  timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu --fbs-kernel $k > gpurun_out/b_c2_$k.json 2> gpurun_out/b_c2_$k.err || tail -3 gpurun_out/b_c2_$k.err
  python - <<PY
import json;d=json.load(open('gpurun_out/b_c2_$k.json'));print('c2 kernel=$k','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4), d['roofline']['bound'], d['roofline'].get('peak'), d['config']['fbs_kernel'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_fbs_jit -c 1 -s 4 -o gpurun_out/b_jit_c2 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu1.log 2>&1; tail -2 gpurun_out/b_ncu1.log
