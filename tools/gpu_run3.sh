#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -120 > gpurun_out/pytest3.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest3.log
timeout 600 python tools/enc_bench.py 256 > gpurun_out/enc_bench3.log 2>&1
for v in "SSDK_ENC_DEBUG=1" "SSDK_ENC_TPC=4" "SSDK_ENC_TPC=8" "SSDK_ENC_DEBUG=1 SSDK_ENC_TPC=8"; do
  echo "== $v" >> gpurun_out/enc_bench3.log
  env $v timeout 300 python tools/profile_encode300.py 2>&1 | grep "encode SSD300" >> gpurun_out/enc_bench3.log
done
timeout 300 python tools/profile_loss.py > gpurun_out/loss3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:ssd_loss -c 2 -o gpurun_out/loss_full -f python tools/profile_loss.py > gpurun_out/ncu_loss.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_micro_full -f python tools/profile_encode.py 64 > gpurun_out/ncu_enc_micro.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench3.json 2> gpurun_out/bench3.err
SSDK_NO_DIRECT=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro > gpurun_out/bench3_nodirect.json 2> gpurun_out/bench3_nodirect.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step3.csv python tools/profile_step.py step > gpurun_out/profile_step3.log 2>&1
tail -6 gpurun_out/pytest3.log; cat gpurun_out/enc_bench3.log | tail -22; cat gpurun_out/loss3.log | tail -3; tail -c 600 gpurun_out/bench3.json; tail -c 300 gpurun_out/bench3_nodirect.json
