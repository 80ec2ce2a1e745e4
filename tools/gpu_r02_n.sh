#!/bin/bash
# Round-2 GPU session N (2 GPUs): new general-kernel tests, then the driver's 2-GPU launch line of bench.py
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session N" > gpurun_out/n_log.txt
timeout 600 python -m pytest tests/test_gpu_general_decoder.py -q -k "thirty_two or wide_beam" 2>&1 | tail -15 >> gpurun_out/n_log.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/n_scale2.json 2> gpurun_out/n_scale2.err
echo "torchrun rc=$?" >> gpurun_out/n_log.txt
python - >> gpurun_out/n_log.txt <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/n_scale2.json").read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value %.0f ms/step %.2f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/n_scale2.err >> gpurun_out/n_log.txt
