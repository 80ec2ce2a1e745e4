#!/bin/bash
# Round-2 GPU session V: libm tables of the decoder's log_sum_exp chains in shared memory -- parity, then bench
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session V" > gpurun_out/v_log.txt
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_headline.py -q -x -k "not transcripts and not am_" 2>&1 | tail -4 >> gpurun_out/v_log.txt
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/v_bench$i.json 2>> gpurun_out/v_err.txt
python - gpurun_out/v_bench$i.json >> gpurun_out/v_log.txt <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print("ms/step %.2f decode %.2f e2e %.2f" % (d["ms_per_step"], d["stages_ms"]["decode"], d["e2e"]["ms_per_step"]), d["roofline_all"]["decode"]["phase_share"], d["clocks"]["reasons"])
PY
done
