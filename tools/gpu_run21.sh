#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -40 > gpurun_out/pytest21.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest21.log
timeout 240 python tools/enc_bench.py 256 > gpurun_out/enc_bench21.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench21.json 2> gpurun_out/bench21.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step21.csv python tools/profile_step.py step > gpurun_out/profile_step21.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_enc21.csv python tools/profile_encode.py 256 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke21.log 2>&1
tail -5 gpurun_out/pytest21.log; cat gpurun_out/enc_bench21.log | tail -14; tail -c 400 gpurun_out/bench21.json; tail -2 gpurun_out/smoke21.log; grep -c enc_ gpurun_out/launches_enc21.csv
