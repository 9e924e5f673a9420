#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r02_topo_n$N.txt
for flag in "" "--no-numa"; do
tag=$( [ -z "$flag" ] && echo numa || echo nonuma )
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload c4 --sub none --no-cpu $flag > gpurun_out/r02_n${N}_c4_$tag.json 2> gpurun_out/r02_n${N}_c4_$tag.err || tail -3 gpurun_out/r02_n${N}_c4_$tag.err
done
python - <<PY
import json
for t in ('numa','nonuma'):
    try:
        d=json.load(open('gpurun_out/r02_n${N}_c4_%s.json'%t))
        print(t,'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'], d['e2e'].get('pcie_gbs_per_rank'), d['config'].get('host_affinity'))
    except Exception as e: print(t,'ERR',e)
PY
cat gpurun_out/r02_topo_n$N.txt | cut -c1-150
