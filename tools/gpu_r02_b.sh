#!/bin/bash
# Round-2 GPU session B: headline parity tests (all of them, no -x), decoder tests, then two full bench lines (exact / MUFU h path)
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== headline + decoder tests" > gpurun_out/b_log.txt
timeout 1500 python -m pytest tests/test_gpu_headline.py tests/test_gpu_decoder.py -q -s 2>&1 | grep -v "TensorFlow: none\|Coqui STT:" | tail -80 >> gpurun_out/b_log.txt
echo "rc=$?" >> gpurun_out/b_log.txt
for h in 1 0; do
  echo "== bench EXACT_H=$h (with CPU baseline + parity)" >> gpurun_out/b_log.txt
  STT_B200_LSTM_EXACT_H=$h timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench_h$h.json 2>> gpurun_out/b_bench_err.txt
  echo "rc=$?" >> gpurun_out/b_log.txt
  python - $h >> gpurun_out/b_log.txt <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/b_bench_h%s.json"%sys.argv[1]))
    print("value %.0f e2e %.0f ms/step %.2f e2e_ms %.2f stages %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stages_ms"].items()}))
    print("parity", json.dumps(d.get("parity")))
    cb=d.get("cpu_baseline",{})
    print("cpu", cb.get("value"), cb.get("repeat_value"), cb.get("cores"), cb.get("error"), cb.get("trace"))
except Exception as e:
    print("parse failed", e)
PY
done
