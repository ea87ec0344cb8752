#!/bin/bash
# 2-GPU run: the data-parallel test and the bench under torchrun (pipelined e2e with the all-gather, train step, fixed-global-batch
# mode, NCCL check); before that, on one GPU: the new stand-alone ops tests and the training check of case 3 on three input seeds
# with and without the fused weight operand
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for seed in 11 12 13; do for f in 1 0; do
  SSDK_CHECK_SEED=$seed SSDK_FUSE_B=$f timeout 200 python tools/train_check.py --case 3 2>&1 | grep "c1/\|c2/\|weights after\|CASE" | sed "s/^/seed $seed fuse $f: /" >> gpurun_out/train_case3_seeds.log
done; done
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_train_dist.py "tests/test_gpu_train.py::test_small_graph_gradients" -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_n2.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_n2.log
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench exit $?" >> gpurun_out/pytest_n2.log
grep -h "NVLS\|via P2P\|Connected all" gpurun_out/nccl_*.log | sort | uniq -c | head -12 > gpurun_out/nccl_summary.txt
rm -f gpurun_out/nccl_*.log
cat gpurun_out/train_case3_seeds.log | cut -c1-170; tail -8 gpurun_out/pytest_n2.log; tail -c 3500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
