#!/bin/bash
# Round-2 GPU session S: LSTM kernels for batches <= 128 (A/B), clock sampling of the bench line
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session S" > gpurun_out/s_log.txt
timeout 600 python tools/lstm_small_check.py 2>&1 | grep -v "Coqui\|TensorFlow" >> gpurun_out/s_log.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s_bench.json 2>> gpurun_out/s_err.txt
python -c "
import json
d=json.load(open('gpurun_out/s_bench.json')); print('ms/step %.2f e2e %.2f' % (d['ms_per_step'], d['e2e']['ms_per_step']), d['clocks'])
" >> gpurun_out/s_log.txt
