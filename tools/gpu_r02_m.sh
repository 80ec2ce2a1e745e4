#!/bin/bash
# Round-2 GPU session M: BASELINE configs[1] (one stream in 20 ms chunks) and configs[4] (beam sweep) on the round-2 build
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session M" > gpurun_out/m_log.txt
timeout 600 python tools/stream_latency.py > gpurun_out/m_stream_latency.json 2>> gpurun_out/m_err.txt
echo "stream rc=$?" >> gpurun_out/m_log.txt
for beam in 100 2000; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --beam $beam > gpurun_out/m_beam_$beam.json 2>> gpurun_out/m_err.txt
  echo "beam $beam rc=$?" >> gpurun_out/m_log.txt
done
python - >> gpurun_out/m_log.txt <<'PY'
import json
print(open("gpurun_out/m_stream_latency.json").read()[:1500])
for b in (100, 2000):
    try:
        d=json.load(open("gpurun_out/m_beam_%d.json"%b)); print(b, "ms/step %.2f decode %.2f e2e %.2f" % (d["ms_per_step"], d["stages_ms"]["decode"], d["e2e"]["ms_per_step"]))
    except Exception as e: print(b, "parse failed", e)
PY
tail -3 gpurun_out/m_err.txt >> gpurun_out/m_log.txt
