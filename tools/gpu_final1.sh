#!/bin/bash
# 1-GPU record: parity suite, the driver-style bench line + reference arm, launch lists of all workloads, ncu of the top kernels
tag=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
for w in linear mlp conv convnet; do
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches_${w}.csv python bench.py --workload $w --profile --no-graph --steps 2 --warmup 3 \
    > gpurun_out/${tag}_launches_${w}.log 2>&1
  echo "launches $w rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 3 -o gpurun_out/${tag}_gemm_step -f \
   python bench.py --workload linear --profile --no-graph --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
