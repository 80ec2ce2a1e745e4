#!/bin/bash
# Round-2 GPU session O: ping-pong LSTM with part of each CTA's weight slice resident in shared memory
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session O" > gpurun_out/o_log.txt
for mode in 6 7; do
  STT_B200_VERBOSE=1 timeout 240 python tools/lstm_mode_check.py $mode 2048 256 2>&1 | grep -v "Coqui\|TensorFlow" >> gpurun_out/o_log.txt
  echo "mode $mode rc=$?" >> gpurun_out/o_log.txt
done
