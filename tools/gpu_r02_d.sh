#!/bin/bash
# Round-2 GPU session D: parity of the barrier-reduced decoder, lifecycle fixes, restated transcript gate, bench A/B (eager LM variant)
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== decoder + lifecycle + headline (fast ones first)" > gpurun_out/d_log.txt
timeout 900 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_lifecycle.py tests/test_gpu_ctcdecoder_api.py -q -x 2>&1 | tail -15 >> gpurun_out/d_log.txt
summ() {
python - "$1" >> gpurun_out/d_log.txt <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("value %.0f e2e %.0f ms/step %.2f e2e_ms %.2f stages %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], {k: round(v,2) for k,v in d["stages_ms"].items()}))
    dec=d["roofline_all"]["decode"]
    print({k: dec[k] for k in dec if k not in ("note",)})
except Exception as e:
    print("parse failed", e)
PY
}
echo "== bench default" >> gpurun_out/d_log.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench.json 2>> gpurun_out/d_err.txt
summ gpurun_out/d_bench.json
echo "== bench eager-LM variant" >> gpurun_out/d_log.txt
STT_B200_LIB=$PWD/build/libstt_b200_eager.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/d_bench_eager.json 2>> gpurun_out/d_err.txt
summ gpurun_out/d_bench_eager.json
echo "== headline tests" >> gpurun_out/d_log.txt
timeout 1500 python -m pytest tests/test_gpu_headline.py -q -s 2>&1 | grep -v "TensorFlow: none\|Coqui STT:" | tail -60 >> gpurun_out/d_log.txt
