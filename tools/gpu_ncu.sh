#!/bin/bash
# One 1-GPU call: parity suite, quick per-workload step times, then `ncu --set full` captures of the dominant kernels.
#   bash tools/gpu_ncu.sh <tag> [tests] [times] [ncu]
tag=${1:-r02}; shift
what=${*:-tests times ncu}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
if [[ $what == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
fi
if [[ $what == *times* ]]; then
  for w in linear mlp conv convnet; do
    timeout 300 python bench.py --workload $w --profile --steps 30 --warmup 5 2>&1 | tail -1
  done | tee gpurun_out/${tag}_times.log
fi
if [[ $what == *ncu* ]]; then
  timeout 600 $NCU -k regex:"conv_bwd_fused|conv_fwd_tz" -c 2 -o gpurun_out/${tag}_conv_step -f \
     python bench.py --workload conv --profile --no-graph --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
  timeout 600 $NCU -k regex:conv_bwd_fused -c 1 -o gpurun_out/${tag}_conv_bwd_general -f \
     python tools/conv_probe.py bwd 256 3 224 224 64 3 3 > gpurun_out/${tag}_ncu_convg.log 2>&1; echo "ncu conv general rc=$?"
  timeout 600 $NCU -k regex:gemm_tc_kernel -c 3 -o gpurun_out/${tag}_gemm_step -f \
     python bench.py --workload linear --profile --no-graph --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
  ls -la gpurun_out/*.ncu-rep
fi
