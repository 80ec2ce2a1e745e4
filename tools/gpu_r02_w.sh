#!/bin/bash
# Round-2 GPU session W (4 GPUs): the driver's torchrun line of bench.py
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/w_scale4.json 2> gpurun_out/w_scale4.err
echo "torchrun rc=$?" > gpurun_out/w_log.txt
python - >> gpurun_out/w_log.txt <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/w_scale4.json").read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value %.0f ms/step %.2f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/w_scale4.err >> gpurun_out/w_log.txt
