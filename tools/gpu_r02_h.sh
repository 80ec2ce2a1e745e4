#!/bin/bash
# Round-2 GPU session H: CTA-pair GEMM -- numerics (GEMM unit tests, AM vs oracle, headline rows) then A/B bench
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session H" > gpurun_out/h_log.txt
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_mfcc_am.py tests/test_gpu_e2e.py -q -x 2>&1 | tail -15 >> gpurun_out/h_log.txt
run() {
  echo "== $1" >> gpurun_out/h_log.txt
  STT_B200_GEMM_PAIR=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/h_$1.json 2>> gpurun_out/h_err.txt
  python - gpurun_out/h_$1.json >> gpurun_out/h_log.txt <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.2f e2e %.2f stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stages_ms"].items()}))
    print({k: round(v["frac"],3) for k,v in d["roofline_all"].items()}, d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
}
run pair 1
run single 0
run pair2 1
timeout 900 python -m pytest tests/test_gpu_headline.py -q -x -k "am_ or rows_equal or partial" 2>&1 | tail -8 >> gpurun_out/h_log.txt
tail -5 gpurun_out/h_err.txt >> gpurun_out/h_log.txt
