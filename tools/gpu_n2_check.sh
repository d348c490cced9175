#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_pytest_n2.log
for w in linear mlp; do timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_timeline.py $w > gpurun_out/${tag}_timeline_${w}_n2.log 2>&1; grep -v "^\*\*\|OMP_NUM\|Warn\|warn\|^$" gpurun_out/${tag}_timeline_${w}_n2.log | cut -c1-150; done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --workload mlp --others linear,convnet --cpu-budget 1 > gpurun_out/${tag}_mlp_n2.json 2> gpurun_out/${tag}_mlp_n2.err; echo "mlp rc=$?"
python - <<'PY'
import json
for line in open("gpurun_out/TAG_mlp_n2.json".replace("TAG","'"$tag"'".strip("'"))):
    if line.startswith("{"):
        d=json.loads(line); print("mlp n2 ms/step", d["ms_per_step"], d["ms_per_step_stats"], "parity", d.get("exchange_parity"))
        for k,v in (d.get("other_configs") or {}).items(): print(" other", k, v.get("ms_per_step"), v.get("error"))
PY
