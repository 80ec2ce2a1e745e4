#!/bin/bash
# Round-2 GPU session C (re-entry baseline): whole GPU suite, production bench line, ncu launch list, full captures of decoder + LSTM.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/c_gpu.txt 2>&1
echo "== full gpu suite" > gpurun_out/c_log.txt
timeout 1800 python -m pytest tests -m gpu -q -s --durations=15 2>&1 | grep -v "TensorFlow: none\|Coqui STT:" | tail -80 >> gpurun_out/c_log.txt
echo "rc=$?" >> gpurun_out/c_log.txt
echo "== bench" >> gpurun_out/c_log.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c_bench.json 2>> gpurun_out/c_bench_err.txt
echo "rc=$?" >> gpurun_out/c_log.txt
echo "== ncu launch list" >> gpurun_out/c_log.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_bench.log 2>&1
echo "ncu rc=$?" >> gpurun_out/c_log.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:decoder_step_kernel -c 1 -o gpurun_out/c_decoder \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_dec.log 2>&1
echo "ncu dec rc=$?" >> gpurun_out/c_log.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lstm_pp_kernel -c 1 -o gpurun_out/c_lstm \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu_lstm.log 2>&1
echo "ncu lstm rc=$?" >> gpurun_out/c_log.txt
ls -la gpurun_out >> gpurun_out/c_log.txt
