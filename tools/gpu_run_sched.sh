#!/bin/bash
# two-stream schedule + pipelined host API: parity tests, then same-box A/B of the schedule knobs (short bench runs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedule.py tests/test_gpu_model.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/pytest_sched.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_sched.log
for cfg in "SSDK_OVERLAP=0" "SSDK_OVERLAP=1" "SSDK_OVERLAP_R=14" "SSDK_OVERLAP_R=50" "SSDK_OVERLAP=0" "SSDK_OVERLAP=1"; do
  echo "== $cfg" >> gpurun_out/bench_sched.log
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-micro 2>> gpurun_out/bench_sched.err | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    print(json.dumps({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches')} | {'e2e': j['e2e']['value'], 'e2e_ms': j['e2e']['ms_per_step'], 'conv_ms': j['roofline']['conv_ms_per_step'], 'clk': j['clocks']}))
" >> gpurun_out/bench_sched.log
done
tail -8 gpurun_out/pytest_sched.log; cat gpurun_out/bench_sched.log; tail -5 gpurun_out/bench_sched.err
