#!/bin/bash
# Round-2 GPU session I: the multicast LSTM (cluster of 8) -- bit-identical to the default? faster?
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session I" > gpurun_out/i_log.txt
STT_B200_VERBOSE=1 timeout 240 python tools/lstm_mode_check.py 9 2048 256 >> gpurun_out/i_log.txt 2>&1
echo "rc=$?" >> gpurun_out/i_log.txt
nvidia-smi --query-gpu=name,utilization.gpu --format=csv >> gpurun_out/i_log.txt 2>&1
