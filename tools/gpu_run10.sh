#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SSDK_ENC_DEBUG=2 timeout 300 python tools/profile_encode.py 256 2>&1 | grep "enc matching" | tail -2 > gpurun_out/enc_prof10.log
SSDK_ENC_DEBUG=2 timeout 300 python tools/profile_encode300.py 2>&1 | grep "enc matching" | tail -2 >> gpurun_out/enc_prof10.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_reference_goldens.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest10.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro > gpurun_out/bench10.json 2> gpurun_out/bench10.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step10.csv python tools/profile_step.py step > gpurun_out/profile_step10.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'conv_first' -c 1 -o gpurun_out/conv_first_full -f python tools/profile_step.py step > gpurun_out/ncu_conv_first.log 2>&1
cat gpurun_out/enc_prof10.log; tail -3 gpurun_out/pytest10.log; tail -c 400 gpurun_out/bench10.json
