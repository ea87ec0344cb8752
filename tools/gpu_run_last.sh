cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_schedule.py tests/test_gpu_ops.py "tests/test_gpu_model.py::test_ssd300_batch32_layers" -m gpu -q -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/pytest_last.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_last.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-micro > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err
cat gpurun_out/pytest_last.log; head -c 400 gpurun_out/bench_last.json
