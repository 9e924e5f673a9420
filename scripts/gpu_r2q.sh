#!/bin/bash
mkdir -p gpurun_out
T=r02q
timeout 900 python -m pytest tests/test_extras_gpu.py tests/test_study_gpu.py tests/test_fbs_gpu.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/${T}_tests.log
timeout 300 python scripts/extras_probe.py > gpurun_out/${T}_extras_probe.txt 2>&1; tail -12 gpurun_out/${T}_extras_probe.txt
