#!/bin/bash
# GPU session: FBS parity tests, A/B of the FBS kernels on C2, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fbs_gpu.py -x -q -m gpu > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/a_tests.log
for k in 0 3; do
  timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu --fbs-kernel $k > gpurun_out/a_c2_$k.json 2> gpurun_out/a_c2_$k.err || tail -3 gpurun_out/a_c2_$k.err
  python - <<PY
import json;d=json.load(open('gpurun_out/a_c2_$k.json'));print('c2 kernel=$k','value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],'frac',round(d['roofline']['frac'],3),'kernel_ms',round(d['roofline']['kernel_ms'],4), d['roofline']['bound'], d['roofline'].get('peak'))
PY
done
GP_KERNELS=3,4 GP_SIZES=320,4736,28416,65536,262144 timeout 300 python scripts/latency_probe.py > gpurun_out/a_lat.log 2>&1; cat gpurun_out/a_lat.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbs_onchip -c 1 -s 4 -o gpurun_out/a_onchip_c2 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/a_ncu1.log 2>&1; tail -2 gpurun_out/a_ncu1.log
