#!/bin/bash
# Round-2 GPU session F: 8-byte candidate records (5632 candidates in shared memory) vs the previous build
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session F" > gpurun_out/f_log.txt
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_headline.py -q -x -k "not transcripts" 2>&1 | tail -5 >> gpurun_out/f_log.txt
run() {
  echo "== $1 batch $2" >> gpurun_out/f_log.txt
  STT_B200_LIB=$3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $2 > gpurun_out/f_$1_$2.json 2>> gpurun_out/f_err.txt
  python - gpurun_out/f_$1_$2.json >> gpurun_out/f_log.txt <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.2f e2e %.2f stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stages_ms"].items()}))
    dec=d["roofline_all"]["decode"]; print({k: dec[k] for k in dec if k not in ("note",)})
except Exception as e:
    print("parse failed", e)
PY
}
run new 256 ""
run prev 256 $PWD/build/libstt_b200_prev.so
run new 256 ""
