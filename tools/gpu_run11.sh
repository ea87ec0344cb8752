#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_codec.py tests/test_gpu_batch_assembly.py tests/test_gpu_reference_goldens.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/pytest11.log
timeout 600 python tools/enc_bench.py 256 > gpurun_out/enc_bench11.log 2>&1
SSDK_ENC_DEBUG=2 timeout 300 python tools/profile_encode.py 256 2>&1 | grep "enc matching" | tail -1 >> gpurun_out/enc_bench11.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_tiles -c 1 -o gpurun_out/enc_micro_full -f python tools/profile_encode.py 64 > gpurun_out/ncu_enc_micro.log 2>&1
tail -4 gpurun_out/pytest11.log; cat gpurun_out/enc_bench11.log | tail -16
