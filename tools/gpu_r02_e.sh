#!/bin/bash
# Round-2 GPU session E: is the decoder held back by its 64-register cap / by sharing an SM?  decode time at 128 utterances
# (one CTA per SM) with the production build (64 registers) and a 128-register build, and at 256 (two per SM).
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session E" > gpurun_out/e_log.txt
run() {
  echo "== $1 batch $2" >> gpurun_out/e_log.txt
  STT_B200_LIB=$3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --batch $2 > gpurun_out/e_$1_$2.json 2>> gpurun_out/e_err.txt
  python - gpurun_out/e_$1_$2.json >> gpurun_out/e_log.txt <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.2f stages %s" % (d["ms_per_step"], {k: round(v,2) for k,v in d["stages_ms"].items()}))
    dec=d["roofline_all"]["decode"]; print(dec.get("phase_share"))
except Exception as e:
    print("parse failed", e)
PY
}
run default 256 ""
run default 128 ""
run regs128 128 $PWD/build/libstt_b200_regs128.so
run default 148 ""
run regs128 148 $PWD/build/libstt_b200_regs128.so
