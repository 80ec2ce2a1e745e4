#!/bin/bash
# Round-2 final check: the driver's GPU gates on the final tree -- whole GPU suite, smoke(), default bench line
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== final" > gpurun_out/final_log.txt
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 >> gpurun_out/final_log.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> gpurun_out/final_log.txt
timeout 900 python bench.py > gpurun_out/final_bench.json 2>> gpurun_out/final_err.txt
echo "bench rc=$?" >> gpurun_out/final_log.txt
python - >> gpurun_out/final_log.txt <<'PY'
import json
d=json.load(open("gpurun_out/final_bench.json"))
print("value %.0f e2e %.0f ms/step %.2f e2e_ms %.2f steps %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["steps"]))
print({k: round(v,2) for k,v in d["stages_ms"].items()}); print(d["roofline"]["frac"], d["clocks"], d["cpu_baseline"].get("value"), d["parity"]["decoder_identical_to_reference_on_gpu_probs"])
PY
