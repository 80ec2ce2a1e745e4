#!/bin/bash
# Round-2 GPU session G: the general decoder kernel (pruning, wide alphabets, UTF-8 bytes mode) against the reference
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session G" > gpurun_out/g_log.txt
timeout 1200 python -m pytest tests/test_gpu_general_decoder.py tests/test_gpu_ctcdecoder_api.py -q 2>&1 | tail -120 >> gpurun_out/g_log.txt
