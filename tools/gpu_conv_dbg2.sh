#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "conv" > gpurun_out/${1}_pytest_conv.log 2>&1; echo "pytest conv rc=$?"; tail -3 gpurun_out/${1}_pytest_conv.log
for d in 0 32 3 35; do
  echo -n "uniform dbg=$d  "; NK_CONV_DBG=$d timeout 120 python bench.py --workload conv --profile --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])"
done 2>&1 | tee gpurun_out/${1}_conv_dbg.log
for d in 0 32; do
  echo -n "general dbg=$d  "; NK_CONV_DBG=$d timeout 120 python tools/conv_probe.py bwd 256 3 224 224 64 3 3 20 2>&1 | tail -1 | cut -c1-300
done 2>&1 | tee -a gpurun_out/${1}_conv_dbg.log
