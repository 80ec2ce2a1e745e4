#!/bin/bash
# Round-2 GPU session R: 128-register decoder build for batches that fit one CTA per SM -- parity, stream latency, small-batch bench
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session R" > gpurun_out/r_log.txt
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_headline.py tests/test_gpu_e2e.py tests/test_gpu_lifecycle.py -q -x -k "not transcripts and not am_" 2>&1 | tail -6 >> gpurun_out/r_log.txt
timeout 600 python tools/stream_latency.py > gpurun_out/r_stream_latency.json 2>> gpurun_out/r_err.txt
cat gpurun_out/r_stream_latency.json >> gpurun_out/r_log.txt
for bsz in 128 256; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $bsz > gpurun_out/r_bench_$bsz.json 2>> gpurun_out/r_err.txt
  python - gpurun_out/r_bench_$bsz.json >> gpurun_out/r_log.txt <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], "ms/step %.2f decode %.2f e2e %.2f" % (d["ms_per_step"], d["stages_ms"]["decode"], d["e2e"]["ms_per_step"]))
PY
done
