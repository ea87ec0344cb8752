cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_keras_fit.py tests/test_gpu_schedule.py::test_predict_stream_matches_per_batch_calls -m gpu -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/pytest_last.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_last.log
cat gpurun_out/pytest_last.log
