#!/bin/bash
# Round-2 GPU session K: whole GPU suite on the current build, production bench line (with CPU arm + parity), ncu launch list,
# full captures of the decoder, the ping-pong LSTM and the CTA-pair GEMM
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/k_gpu.txt 2>&1
echo "== full gpu suite" > gpurun_out/k_log.txt
timeout 2400 python -m pytest tests -m gpu -q -s --durations=10 2>&1 | grep -v "TensorFlow: none\|Coqui STT:" | tail -60 >> gpurun_out/k_log.txt
echo "== bench" >> gpurun_out/k_log.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/k_bench.json 2>> gpurun_out/k_bench_err.txt
echo "rc=$?" >> gpurun_out/k_log.txt
echo "== ncu launch list" >> gpurun_out/k_log.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/k_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/k_ncu_bench.log 2>&1
echo "ncu rc=$?" >> gpurun_out/k_log.txt
for k in decoder_step_kernel lstm_pp_kernel gemm2_tc_kernel; do
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -o gpurun_out/k_$k \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/k_ncu_$k.log 2>&1
  echo "ncu $k rc=$?" >> gpurun_out/k_log.txt
done
ls -la gpurun_out >> gpurun_out/k_log.txt
