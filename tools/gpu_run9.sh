#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -100 > gpurun_out/pytest9.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest9.log
timeout 600 python tools/enc_bench.py 256 > gpurun_out/enc_bench9.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench9.json 2> gpurun_out/bench9.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step9.csv python tools/profile_step.py step > gpurun_out/profile_step9.log 2>&1
# full captures of the non-conv kernels of one step (one launch each) and of the encoder
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'conv_first|nms_kernel|maxpool|topk|dec_prepare|l2norm|preprocess|im2col' -c 14 -o gpurun_out/step_small_full -f python tools/profile_step.py step > gpurun_out/ncu_step_small.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_micro_full -f python tools/profile_encode.py 64 > gpurun_out/ncu_enc_micro.log 2>&1
tail -6 gpurun_out/pytest9.log; cat gpurun_out/enc_bench9.log | tail -18; tail -c 600 gpurun_out/bench9.json
