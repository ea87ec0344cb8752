#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -40 > gpurun_out/pytest_verify.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_verify.log
timeout 240 python tools/enc_bench.py 256 > gpurun_out/enc_bench_verify.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_verify.json 2> gpurun_out/bench_verify.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step_verify.csv python tools/profile_step.py step > gpurun_out/profile_step_verify.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_enc_verify.csv python tools/profile_encode.py 256 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_verify.log 2>&1
tail -5 gpurun_out/pytest_verify.log; cat gpurun_out/enc_bench_verify.log | tail -14; tail -c 400 gpurun_out/bench_verify.json; tail -2 gpurun_out/smoke_verify.log; grep -c enc_ gpurun_out/launches_enc_verify.csv
