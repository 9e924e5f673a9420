#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nr3_solve -c 1 -s 4 -o gpurun_out/h_nr3_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu > gpurun_out/h_ncu1.log 2>&1; tail -2 gpurun_out/h_ncu1.log
