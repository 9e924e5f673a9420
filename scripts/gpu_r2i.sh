#!/bin/bash
# source-level ncu capture of the NR3 phase-lane kernel on C3 (stall samples per line)
mkdir -p gpurun_out
T=r02i
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nr3_phase_kernel -c 1 -s 3 -o gpurun_out/${T}_nr3_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu1.log 2>&1; tail -2 gpurun_out/${T}_ncu1.log
