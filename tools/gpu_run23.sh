#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -40 > gpurun_out/pytest23.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest23.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench23.json 2> gpurun_out/bench23.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step23.csv python tools/profile_step.py step > gpurun_out/profile_step23.log 2>&1
tail -5 gpurun_out/pytest23.log; tail -c 500 gpurun_out/bench23.json
