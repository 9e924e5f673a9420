#!/bin/bash
# C4 e2e re-check: link rate of the box, then the default host path twice and with 1 / 3 chunks
mkdir -p gpurun_out
T=r02u
timeout 120 python scripts/pcie_probe.py > gpurun_out/${T}_pcie.txt 2>&1; tail -5 gpurun_out/${T}_pcie.txt
for v in a b; do timeout 300 python bench.py --workload c4 --no-cpu --steps 20 > gpurun_out/${T}_c4_$v.json 2> gpurun_out/${T}_c4_$v.err; done
for c in 1 3; do timeout 300 python bench.py --workload c4 --no-cpu --steps 20 --host-chunks $c > gpurun_out/${T}_c4_chunks$c.json 2> gpurun_out/${T}_c4_chunks$c.err; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_c4_*.json')):
    try:
        d=json.load(open(f)); e=d['e2e']
        print(f.split('/')[-1],'value %.3e'%d['value'],'e2e %.3e'%e['value'],'pcie',round(e['pcie_gbs_per_rank'],1),'vmag-only %.3e'%e['returning_vmag_and_losses_only']['value_this_rank'])
    except Exception as ex: print(f,'ERR',ex)
PY
