#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_codec.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -60 > gpurun_out/pytest2_codec.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest2_codec.log
timeout 600 python tools/enc_bench.py 256 > gpurun_out/enc_bench2.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py --ignore tests/test_gpu_codec.py 2>&1 | tail -80 > gpurun_out/pytest2_rest.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest2_rest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_micro_full -f python tools/profile_encode.py 64 > gpurun_out/ncu_enc_micro.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_ssd300_full -f python tools/profile_encode300.py > gpurun_out/ncu_enc300.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err
tail -4 gpurun_out/pytest2_codec.log; tail -4 gpurun_out/pytest2_rest.log; cat gpurun_out/enc_bench2.log | tail -12; tail -c 1200 gpurun_out/bench2.json
