#!/bin/bash
# Round-2 GPU session L: can the next batch's GEMMs start in the SM space the decoder leaves free?  3-stage ring (fits beside a
# decoder CTA) and chunked tile assignment (the block scheduler places pairs wherever there is room)
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session L" > gpurun_out/l_log.txt
run() {
  echo "== $1" >> gpurun_out/l_log.txt
  env $2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/l_$1.json 2>> gpurun_out/l_err.txt
  python - gpurun_out/l_$1.json >> gpurun_out/l_log.txt <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.2f e2e %.2f stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stages_ms"].items()}), d["clocks"]["reasons"])
except Exception as e:
    print("parse failed", e)
PY
}
run base "X=1"
run s3 "STT_B200_GEMM2_STAGES=3"
run s3c8 "STT_B200_GEMM2_STAGES=3 STT_B200_GEMM2_CHUNK=8"
run s6c8 "STT_B200_GEMM2_CHUNK=8"
run s3c4 "STT_B200_GEMM2_STAGES=3 STT_B200_GEMM2_CHUNK=4"
run base2 "X=1"
run s3c8b "STT_B200_GEMM2_STAGES=3 STT_B200_GEMM2_CHUNK=8"
tail -3 gpurun_out/l_err.txt >> gpurun_out/l_log.txt
