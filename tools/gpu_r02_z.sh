#!/bin/bash
# Round-2 session Z: scorer path at a 300k-word vocabulary (tools/large_vocab_check.py)
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 240 python tools/large_vocab_check.py > gpurun_out/z_large_vocab.json 2> gpurun_out/z_err.txt
echo "rc=$?" > gpurun_out/z_log.txt
tail -5 gpurun_out/z_err.txt >> gpurun_out/z_log.txt
