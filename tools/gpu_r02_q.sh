#!/bin/bash
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
STT_GEN_PROF=1 STT_B200_LIB=$PWD/build/libstt_b200_genprof.so timeout 600 python tools/general_decoder_speed.py 2>/dev/null > gpurun_out/q_prof.json
echo "rc=$?" > gpurun_out/q_log.txt
python -c "
import json
d=json.load(open('gpurun_out/q_prof.json'))
for k,v in d.items(): print(k, round(v['us_per_timestep'],1), v.get('phase_kcycles'), v.get('mean_candidates_per_step'))
" >> gpurun_out/q_log.txt
timeout 900 python -m pytest tests/test_gpu_general_decoder.py tests/test_gpu_ctcdecoder_api.py -q 2>&1 | tail -12 >> gpurun_out/q_log.txt
