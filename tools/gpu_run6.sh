#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -100 > gpurun_out/pytest6.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest6.log
SSDK_LOSS_TIMES=1 timeout 300 python tools/profile_loss.py 2>&1 | tail -6 > gpurun_out/loss6.log
timeout 300 python tools/profile_loss.py >> gpurun_out/loss6.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench6.json 2> gpurun_out/bench6.err
SSDK_NO_FIRST_TC=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro > gpurun_out/bench6_nofirst.json 2> gpurun_out/bench6_nofirst.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step6.csv python tools/profile_step.py step > gpurun_out/profile_step6.log 2>&1
tail -6 gpurun_out/pytest6.log; tail -4 gpurun_out/loss6.log; tail -c 600 gpurun_out/bench6.json; tail -c 300 gpurun_out/bench6_nofirst.json
