#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -60 > gpurun_out/pytest12.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest12.log
timeout 600 python tools/enc_bench.py 256 > gpurun_out/enc_bench12.log 2>&1
SSDK_ENC_DEBUG=2 timeout 300 python tools/profile_encode.py 256 2>&1 | grep "enc matching" | tail -1 >> gpurun_out/enc_bench12.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench12.json 2> gpurun_out/bench12.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step12.csv python tools/profile_step.py step > gpurun_out/profile_step12.log 2>&1
tail -6 gpurun_out/pytest12.log; cat gpurun_out/enc_bench12.log | tail -16; tail -c 500 gpurun_out/bench12.json
