#!/bin/bash
# Multi-GPU call: the N-rank parity test of the fused exchange, then the driver's exact scaling command, then the MLP
# (config 4, strong scaling) as the main workload.   bash tools/gpu_scale.sh <tag> <N>
tag=${1:-r02}; N=${2:-2}
mkdir -p gpurun_out
NK_DP_TEST_WORLD=$N timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest_n${N}.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_n${N}.log
NK_DP_TEST_WORLD=$N timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tests/dp_worker.py > gpurun_out/${tag}_dp_worker_n${N}.log 2>&1; echo "dp_worker rc=$?"; tail -2 gpurun_out/${tag}_dp_worker_n${N}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_scale_n${N}.json 2> gpurun_out/${tag}_scale_n${N}.err; echo "bench rc=$?"
cat gpurun_out/${tag}_scale_n${N}.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 100 --warmup 5 --workload mlp --others none --cpu-budget 1 > gpurun_out/${tag}_mlp_n${N}.json 2> gpurun_out/${tag}_mlp_n${N}.err; echo "mlp rc=$?"
cat gpurun_out/${tag}_mlp_n${N}.json
