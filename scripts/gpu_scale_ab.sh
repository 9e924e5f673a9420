#!/bin/bash
# N GPUs: copy-engine all-gather spread over 1 / 4 / 7 copy streams, C4 and C5 only
N=${1:-8}; T=${2:-r02}
mkdir -p gpurun_out
for ps in 1 4 7; do for w in c4 c5; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --workload $w --push-streams $ps --no-cpu > gpurun_out/${T}_n${N}_${w}_ps$ps.json 2> gpurun_out/${T}_n${N}_${w}_ps$ps.err || tail -3 gpurun_out/${T}_n${N}_${w}_ps$ps.err
done; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_n${N}_c*_ps*.json')):
    try:
        d=json.load(open(f)); r=d['roofline']
        print(f.split('/')[-1],'value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],'kernel_ms',round(r['kernel_ms'],4), round(r['nvlink']['achieved_gbs'],1))
    except Exception as e: print(f,'ERR',e)
PY
