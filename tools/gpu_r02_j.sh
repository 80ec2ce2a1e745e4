#!/bin/bash
# Round-2 GPU session J: 256-thread decoder (two prefixes per thread, 128 registers) -- parity, then A/B; LSTM multicast A/B
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session J" > gpurun_out/j_log.txt
STT_B200_DEC_THREADS=256 timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_headline.py -q -x -k "not transcripts and not am_" 2>&1 | tail -8 >> gpurun_out/j_log.txt
run() {
  echo "== $1" >> gpurun_out/j_log.txt
  env $2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/j_$1.json 2>> gpurun_out/j_err.txt
  python - gpurun_out/j_$1.json >> gpurun_out/j_log.txt <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.2f e2e %.2f stages %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stages_ms"].items()}), d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
}
run base "X=1"
run dec256 "STT_B200_DEC_THREADS=256"
run lstm5 "STT_B200_LSTM_PINGPONG=5"
run base2 "X=1"
run dec256b "STT_B200_DEC_THREADS=256"
run lstm5b "STT_B200_LSTM_PINGPONG=5"
tail -3 gpurun_out/j_err.txt >> gpurun_out/j_log.txt
