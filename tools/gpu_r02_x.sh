#!/bin/bash
# Round-2 GPU session X: e2e pipeline depth 1 / 2 / 3
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session X" > gpurun_out/x_log.txt
for d in 2 3 1 2 3; do
  STT_BENCH_E2E_DEPTH=$d timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/x_d$d.json 2>> gpurun_out/x_err.txt
  python -c "
import json
d=json.load(open('gpurun_out/x_d$d.json')); print('depth $d', 'ms/step %.2f e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']), d['clocks']['reasons'])
" >> gpurun_out/x_log.txt
done
