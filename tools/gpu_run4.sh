#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -100 > gpurun_out/pytest4.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest4.log
timeout 600 python tools/enc_bench.py 256 > gpurun_out/enc_bench4.log 2>&1
for v in "SSDK_ENC_LB_MIN=1" "SSDK_ENC_LB_MIN=1 SSDK_ENC_DEBUG=1"; do
  echo "== $v" >> gpurun_out/enc_bench4.log
  env $v timeout 300 python tools/profile_encode300.py 2>&1 | grep "encode SSD300" >> gpurun_out/enc_bench4.log
done
SSDK_LOSS_TIMES=1 timeout 300 python tools/profile_loss.py > gpurun_out/loss4.log 2>&1
timeout 300 python tools/profile_loss.py >> gpurun_out/loss4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_micro_full -f python tools/profile_encode.py 64 > gpurun_out/ncu_enc_micro.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench4.json 2> gpurun_out/bench4.err
SSDK_NO_HEAD_FUSION=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro > gpurun_out/bench4_nofuse.json 2> gpurun_out/bench4_nofuse.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step4.csv python tools/profile_step.py step > gpurun_out/profile_step4.log 2>&1
tail -6 gpurun_out/pytest4.log; cat gpurun_out/enc_bench4.log | tail -16; grep -E "phases|loss fwd" gpurun_out/loss4.log | tail -8; tail -c 500 gpurun_out/bench4.json; tail -c 300 gpurun_out/bench4_nofuse.json
