#!/bin/bash
# Round-2 GPU session T: pair LSTM kernel as the default for batches <= 128 -- tests that use small batches / streams, stream latency
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== session T" > gpurun_out/t_log.txt
timeout 1200 python -m pytest tests/test_gpu_mfcc_am.py tests/test_gpu_e2e.py tests/test_gpu_lifecycle.py tests/test_gpu_cli.py tests/test_gpu_tflite.py tests/test_gpu_headline.py -q -x -k "not transcripts" 2>&1 | tail -6 >> gpurun_out/t_log.txt
timeout 600 python tools/stream_latency.py > gpurun_out/t_stream_latency.json 2>> gpurun_out/t_err.txt
cat gpurun_out/t_stream_latency.json >> gpurun_out/t_log.txt
STT_B200_LSTM_SMALL_PP=0 timeout 600 python tools/stream_latency.py 2>> gpurun_out/t_err.txt >> gpurun_out/t_log.txt
