#!/bin/bash
# 2-GPU run: the data-parallel tests and the bench under torchrun (train step, fixed-global-batch mode, NCCL check)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_train_dist.py tests/test_gpu_train.py -m gpu -q -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_n2.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_n2.log
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_%p.log timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench exit $?" >> gpurun_out/pytest_n2.log
grep -h "NVLS\|via P2P\|Connected all" gpurun_out/nccl_*.log | sort | uniq -c | head -12 > gpurun_out/nccl_summary.txt
rm -f gpurun_out/nccl_*.log
tail -8 gpurun_out/pytest_n2.log; tail -c 3000 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
