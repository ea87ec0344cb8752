#!/bin/bash
# epilogue variants (SSDK_EPI_PIPE): parity tests with the default, then same-box A/B (short bench runs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/bench_sched.log gpurun_out/bench_sched.err
T="tests/test_gpu_schedule.py tests/test_gpu_model.py tests/test_gpu_reference_goldens.py tests/test_gpu_train.py tests/test_gpu_ops.py"
timeout 1200 python -m pytest $T -m gpu -q -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/pytest_sched.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_sched.log
for cfg in "SSDK_EPI_PIPE=0" "SSDK_EPI_PIPE=2" "SSDK_EPI_PIPE=1" "SSDK_EPI_PIPE=0" "SSDK_EPI_PIPE=2" "SSDK_EPI_PIPE=1"; do
  echo "== $cfg" >> gpurun_out/bench_sched.log
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-micro 2>> gpurun_out/bench_sched.err | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    print(json.dumps({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches')} | {'e2e': j['e2e']['value'], 'e2e_ms': j['e2e']['ms_per_step'], 'conv_ms': j['roofline']['conv_ms_per_step'], 'frac': j['roofline']['frac'], 'clk': j['clocks']['sm_mhz'], 'why': j['clocks']['reasons']}))
" >> gpurun_out/bench_sched.log
done
tail -5 gpurun_out/pytest_sched.log; cat gpurun_out/bench_sched.log; tail -5 gpurun_out/bench_sched.err
