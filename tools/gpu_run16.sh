#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_batch_assembly.py tests/test_gpu_reference_goldens.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest16.log
timeout 240 python tools/enc_bench.py 256 > gpurun_out/enc_bench16.log 2>&1
echo "enc_bench exit $?" >> gpurun_out/enc_bench16.log
SSDK_ENC_DEBUG=2 timeout 300 python tools/profile_encode.py 256 2>&1 | grep "enc matching" | tail -1 >> gpurun_out/enc_bench16.log
tail -3 gpurun_out/pytest16.log; cat gpurun_out/enc_bench16.log | tail -17
