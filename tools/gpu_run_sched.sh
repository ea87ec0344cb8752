#!/bin/bash
# plan features of this round (two-stream schedule, fused weight operand, shared border): parity tests, the training check with and
# without the fused operand, then same-box A/B of the knobs (short bench runs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/bench_sched.log gpurun_out/bench_sched.err
for f in 1 0; do
  SSDK_FUSE_B=$f timeout 200 python tools/train_check.py --case 3 > gpurun_out/train_case3_fuse$f.log 2>&1
done
T="tests/test_gpu_schedule.py tests/test_gpu_model.py tests/test_gpu_reference_goldens.py tests/test_gpu_train.py"
timeout 1200 python -m pytest $T -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_sched.log
echo "exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_sched.log
for cfg in "SSDK_X=0" "SSDK_BN128_PENALTY=1.03" "SSDK_BN128_PENALTY=1.10" "SSDK_OVERLAP_R=37" "SSDK_X=0" "SSDK_BN128_PENALTY=1.03"; do
  echo "== $cfg" >> gpurun_out/bench_sched.log
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-micro 2>> gpurun_out/bench_sched.err | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    print(json.dumps({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches')} | {'e2e': j['e2e']['value'], 'e2e_ms': j['e2e']['ms_per_step'], 'conv_ms': j['roofline']['conv_ms_per_step'], 'frac': j['roofline']['frac'], 'clk': j['clocks']['sm_mhz'], 'why': j['clocks']['reasons']}))
" >> gpurun_out/bench_sched.log
done
grep -n "BAD\|CASE\|weights after\|loss" gpurun_out/train_case3_fuse1.log | head -20; grep -n "BAD\|CASE" gpurun_out/train_case3_fuse0.log | head; tail -12 gpurun_out/pytest_sched.log; cat gpurun_out/bench_sched.log; tail -5 gpurun_out/bench_sched.err
