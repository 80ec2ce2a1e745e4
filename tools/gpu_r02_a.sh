#!/bin/bash
# Round-2 GPU session A: parity of the reworked decoder first (short timeout), then the whole GPU suite, then A/B bench lines.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
echo "== decoder parity (quick)" > gpurun_out/a_log.txt
timeout 600 python -m pytest tests/test_gpu_decoder.py -x -q 2>&1 | tail -15 >> gpurun_out/a_log.txt
echo "rc=$?" >> gpurun_out/a_log.txt
echo "== full gpu suite" >> gpurun_out/a_log.txt
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -60 >> gpurun_out/a_log.txt
echo "rc=$?" >> gpurun_out/a_log.txt
for cfg in "3 1" "0 1" "1 1" "2 1" "3 0"; do
  set -- $cfg
  echo "== bench DEC_FLAGS=$1 EXACT_H=$2" >> gpurun_out/a_log.txt
  STT_B200_DEC_FLAGS=$1 STT_B200_LSTM_EXACT_H=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline \
      > gpurun_out/a_bench_f$1_h$2.json 2>> gpurun_out/a_bench_err.txt
  echo "rc=$?" >> gpurun_out/a_log.txt
  python - "$1" "$2" >> gpurun_out/a_log.txt <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/a_bench_f%s_h%s.json"%(sys.argv[1],sys.argv[2])))
    print("value %.0f e2e %.0f ms/step %.2f e2e_ms %.2f stages %s pageable %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"],
          {k: round(v,2) for k,v in d["stages_ms"].items()}, d["e2e"].get("pageable_unpipelined")))
    print("phase_share", d["roofline_all"]["decode"].get("phase_share"))
except Exception as e:
    print("parse failed", e)
PY
done
echo "== ncu launch list (production flags)" >> gpurun_out/a_log.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_bench.log 2>&1
echo "ncu rc=$?" >> gpurun_out/a_log.txt
echo "== ncu full capture of the decoder step kernel" >> gpurun_out/a_log.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:decoder_step_kernel -c 1 -o gpurun_out/a_decoder \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu_dec.log 2>&1
echo "ncu rc=$?" >> gpurun_out/a_log.txt
ls -la gpurun_out >> gpurun_out/a_log.txt
