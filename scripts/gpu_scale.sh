#!/bin/bash
# bench.py under torchrun at N GPUs (default workload C4 + sub_results), as the driver launches it
N=${1:-2}; T=${2:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${T}_scale_n$N.json 2> gpurun_out/${T}_scale_n$N.err ) 2>&1 | grep real
echo "rc=$?"; grep "\[bench\]" gpurun_out/${T}_scale_n$N.err | tail -8; tail -3 gpurun_out/${T}_scale_n$N.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${T}_scale_n$N.json')); r=d['roofline']
    print('N=$N value %.3e'%d['value'],'ms',round(d['ms_per_step'],3),'e2e %.3e'%d['e2e']['value'],r['kernel'],'frac',round(r['frac'],3),'kernel_ms',round(r['kernel_ms'],4), r.get('nvlink'))
    for k,v in (d.get('sub_results') or {}).items():
        print('  ',k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_step','kernel','kernel_ms','bound','frac','hbm_frac','error')}, v.get('nvlink',{}).get('achieved_gbs'))
except Exception as e: print('ERR',e)
PY
