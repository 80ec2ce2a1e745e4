#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
STT_B200_VERBOSE=1 timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/u_bench.json 2> gpurun_out/u_err.txt
grep "lstm_pp roles" gpurun_out/u_err.txt | tail -2 > gpurun_out/u_log.txt
python -c "
import json
d=json.load(open('gpurun_out/u_bench.json')); print(d['lstm_cycles_per_launch'], d['stages_ms']['lstm'])
" >> gpurun_out/u_log.txt
