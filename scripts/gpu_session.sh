#!/bin/bash
# Full GPU session: parity tests, smoke, bench lines, ncu launch list and full captures
mkdir -p gpurun_out
T=${1:-r1c}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${T}_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_ref_c2.json 2> gpurun_out/${T}_ref_c2.err; echo "ref rc=$?"
for w in c3 c4 c4ts c5 c4nr3 nr3_37_big nr3_123_big fbs34 fbs37; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err || tail -3 gpurun_out/${T}_bench_$w.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench_*.json'))+['gpurun_out/${T}_ref_c2.json']:
    try:
        d=json.load(open(f))
        r=d.get('roofline',{})
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],r.get('bound'),'frac',round(r.get('frac',0),3),'kernel_ms',round(r.get('kernel_ms',0),4),(d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gp_fbs_jit -c 1 -s 4 -o gpurun_out/${T}_jit_c2 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu1.log 2>&1; tail -1 gpurun_out/${T}_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fbs_solve_kernel -c 1 -s 4 -o gpurun_out/${T}_stream_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu2.log 2>&1; tail -1 gpurun_out/${T}_ncu2.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:nr3_(solve|phase)_kernel' -c 1 -s 4 -o gpurun_out/${T}_nr3_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu3.log 2>&1; tail -1 gpurun_out/${T}_ncu3.log
