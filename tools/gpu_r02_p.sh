#!/bin/bash
# Round-2 GPU session P: speed of the general decoder kernel (for DESIGN.md), new tests
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session P" > gpurun_out/p_log.txt
timeout 600 python tools/general_decoder_speed.py 2>/dev/null > gpurun_out/p_general_speed.json
echo "rc=$?" >> gpurun_out/p_log.txt
cat gpurun_out/p_general_speed.json >> gpurun_out/p_log.txt
timeout 600 python -m pytest tests/test_gpu_general_decoder.py tests/test_gpu_gemm.py -q 2>&1 | tail -5 >> gpurun_out/p_log.txt
