#!/bin/bash
# Two-sweep kernel in depth-first order: no load math on no-load phases, L2 prefetch distance sweep; e2e chunk sweep
mkdir -p gpurun_out
T=r02f
timeout 900 python -m pytest tests/test_fbs_gpu.py tests/test_big_feeders.py tests/test_extras_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
for w in c4 c5; do for pf in 0 1 2 4 8; do
  timeout 600 python bench.py --workload $w --prefetch $pf --steps 10 --no-cpu > gpurun_out/${T}_${w}_pf$pf.json 2> gpurun_out/${T}_${w}_pf$pf.err || tail -3 gpurun_out/${T}_${w}_pf$pf.err
done; done
for hc in 2 8 16; do
  timeout 600 python bench.py --workload c4 --host-chunks $hc --steps 10 --no-cpu > gpurun_out/${T}_c4_hc$hc.json 2> gpurun_out/${T}_c4_hc$hc.err || tail -3 gpurun_out/${T}_c4_hc$hc.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_c*_*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],r['kernel'],r['bound'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4))
    except Exception as e: print(f,'ERR',e)
PY
