#!/bin/bash
# Round-2 session Z2: both CLIs after the WAV-reader change
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 100 python -m pytest tests/test_gpu_cli.py -q -m gpu 2>&1 | tail -5 > gpurun_out/z2_log.txt
