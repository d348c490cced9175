#!/bin/bash
# One gpurun call: GPU parity suite, the default bench line, the reference arm, and the ncu launch lists of every workload.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_round.sh <tag> [tests|bench|launches|all]
tag=${1:-r02}; what=${2:-all}
mkdir -p gpurun_out
if [[ $what == all || $what == tests ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
  tail -3 gpurun_out/${tag}_pytest.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
  timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
  cat gpurun_out/${tag}_bench.json
fi
if [[ $what == all || $what == launches ]]; then
  for w in linear mlp conv convnet; do
    timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/${tag}_launches_${w}.csv python bench.py --workload $w --profile --no-graph --steps 2 --warmup 3 \
      > gpurun_out/${tag}_launches_${w}.log 2>&1
    echo "launches $w rc=$?"
  done
fi
