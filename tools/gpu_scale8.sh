#!/bin/bash
# N-GPU call, kept short (charged N x): exchange parity worker at world N, then the driver's exact scaling command.
tag=${1:-r02}; N=${2:-8}
mkdir -p gpurun_out
NK_DP_TEST_WORLD=$N timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tests/dp_worker.py > gpurun_out/${tag}_dp_worker_n${N}.log 2>&1; echo "dp_worker rc=$?"; tail -2 gpurun_out/${tag}_dp_worker_n${N}.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --cpu-budget 4 > gpurun_out/${tag}_scale_n${N}.json 2> gpurun_out/${tag}_scale_n${N}.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/${tag}_scale_n${N}.json; tail -5 gpurun_out/${tag}_scale_n${N}.err
