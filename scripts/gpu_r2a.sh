#!/bin/bash
# Round 2, first GPU session: parity tests, smoke, the new default bench line (C4 + sub_results), reference arm
mkdir -p gpurun_out
T=${1:-r02a}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${T}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
( time timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -8 gpurun_out/${T}_bench.err
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err ) 2>&1 | grep real; echo "ref rc=$?"; tail -3 gpurun_out/${T}_ref.err
python - <<PY
import json
for f in ('gpurun_out/${T}_bench.json','gpurun_out/${T}_ref.json'):
    try:
        d=json.load(open(f))
        r=d.get('roofline',{})
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],r.get('bound'),'frac',round(r.get('frac',0),3),'kernel_ms',round(r.get('kernel_ms',0),4),json.dumps(d.get('cpu_baseline'))[:600])
        for k,v in (d.get('sub_results') or {}).items():
            print('  ',k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','kernel','kernel_ms','bound','frac','hbm_frac','error')})
    except Exception as e: print(f,'ERR',e)
PY
