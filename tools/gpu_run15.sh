#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -60 > gpurun_out/pytest15.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest15.log
timeout 240 python tools/enc_bench.py 256 > gpurun_out/enc_bench15.log 2>&1
echo "enc_bench exit $?" >> gpurun_out/enc_bench15.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench15.json 2> gpurun_out/bench15.err
tail -6 gpurun_out/pytest15.log; cat gpurun_out/enc_bench15.log | tail -16; tail -c 400 gpurun_out/bench15.json
