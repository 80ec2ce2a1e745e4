#!/bin/bash
# Round-2 session Y: the random-LM decoder tests (orders 2/3/6, word and UTF-8 mode) on the GPU
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== y" > gpurun_out/y_log.txt
timeout 600 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_general_decoder.py -q -m gpu -k "random_lms" 2>&1 | tail -40 >> gpurun_out/y_log.txt
