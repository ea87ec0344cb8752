#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_codec.py tests/test_gpu_batch_assembly.py tests/test_gpu_model.py tests/test_gpu_reference_goldens.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/pytest14.log
timeout 600 python tools/enc_bench.py 256 > gpurun_out/enc_bench14.log 2>&1
SSDK_ENC_DEBUG=2 timeout 300 python tools/profile_encode.py 256 2>&1 | grep "enc matching" | tail -1 >> gpurun_out/enc_bench14.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_micro_full -f python tools/profile_encode.py 64 > gpurun_out/ncu_enc_micro.log 2>&1
SSDK_ENC_DEBUG=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_micro_nofinish_full -f python tools/profile_encode.py 64 > gpurun_out/ncu_enc_micro2.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro > gpurun_out/bench14.json 2> gpurun_out/bench14.err
tail -4 gpurun_out/pytest14.log; cat gpurun_out/enc_bench14.log | tail -16; tail -c 300 gpurun_out/bench14.json
