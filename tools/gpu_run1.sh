#!/bin/bash
# first GPU call of round 2: full GPU test suite, encoder sweep, encoder ncu captures, a short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -150 > gpurun_out/pytest1.log
echo "pytest exit ${PIPESTATUS[0]}" >> gpurun_out/pytest1.log
timeout 600 python tools/enc_bench.py 256 > gpurun_out/enc_bench1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_micro_full -f python tools/profile_encode.py 64 > gpurun_out/ncu_enc_micro.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_ssd300_full -f python tools/profile_encode300.py > gpurun_out/ncu_enc300.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -5 gpurun_out/pytest1.log; cat gpurun_out/enc_bench1.log | tail -12; tail -c 1500 gpurun_out/bench1.json
