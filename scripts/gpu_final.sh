#!/bin/bash
# Round-2 evidence session: full GPU tests, smoke, default bench + reference arm, DMMA probe, ncu launch list and full captures
mkdir -p gpurun_out
T=${1:-r02k}
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${T}_ref.json 2> gpurun_out/${T}_ref.err; echo "ref rc=$?"
timeout 120 python scripts/dmma_probe.py > gpurun_out/${T}_dmma.json 2> gpurun_out/${T}_dmma.err; cat gpurun_out/${T}_dmma.json
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_default_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_l.log 2>&1; tail -1 gpurun_out/${T}_ncu_l.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fbs_solve_kernel -c 1 -s 4 -o gpurun_out/${T}_stream_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu1.log 2>&1; tail -1 gpurun_out/${T}_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_fbs_jit -c 1 -s 4 -o gpurun_out/${T}_jit_c2 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu2.log 2>&1; tail -1 gpurun_out/${T}_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nr3_phase_kernel -c 1 -s 4 -o gpurun_out/${T}_nr3_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu3.log 2>&1; tail -1 gpurun_out/${T}_ncu3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbs_solve_kernel -c 1 -s 4 -o gpurun_out/${T}_stream_c5 python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu4.log 2>&1; tail -1 gpurun_out/${T}_ncu4.log
