#!/bin/bash
# Final round-2 validation of the tree as committed: full GPU tests, smoke, default bench + reference arm
mkdir -p gpurun_out
T=${1:-r02t}
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/${T}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
python - <<PY
import json
for f in ('gpurun_out/${T}_bench.json','gpurun_out/${T}_ref.json'):
    try:
        d=json.load(open(f)); r=d.get('roofline',{})
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],r.get('bound'),'frac',round(r.get('frac',0),3),'kernel_ms',round(r.get('kernel_ms',0),4))
        for k,v in (d.get('sub_results') or {}).items():
            print('  ',k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','kernel','kernel_ms','bound','frac','hbm_frac','error')})
    except Exception as e: print(f,'ERR',e)
PY
