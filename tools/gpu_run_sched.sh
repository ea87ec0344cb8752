#!/bin/bash
# new plan features (two-stream schedule, fused weight operand, shared border, folded L2Normalization): parity tests, a knock-out
# pass if they fail, then same-box A/B of the knobs (short bench runs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/bench_sched.log gpurun_out/bench_sched.err
timeout 300 python tools/e2e_diag.py > gpurun_out/e2e_diag.log 2>&1
T="tests/test_gpu_schedule.py tests/test_gpu_model.py tests/test_gpu_reference_goldens.py tests/test_gpu_train.py"
timeout 1200 python -m pytest $T -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_sched.log
rc=${PIPESTATUS[0]}
echo "exit $rc" >> gpurun_out/pytest_sched.log
if [ "$rc" != "0" ]; then
  for k in SSDK_FUSE_B SSDK_SHARED_BORDER SSDK_FOLD_L2N; do
    env $k=0 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_reference_goldens.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/pytest_sched_$k.log
    echo "== $k=0: exit ${PIPESTATUS[0]}" >> gpurun_out/pytest_sched.log
  done
fi
OFF="SSDK_FUSE_B=0 SSDK_SHARED_BORDER=0 SSDK_FOLD_L2N=0"
for cfg in "$OFF" "SSDK_SHARED_BORDER=0 SSDK_FOLD_L2N=0" "SSDK_FUSE_B=0 SSDK_FOLD_L2N=0" "SSDK_FUSE_B=0 SSDK_SHARED_BORDER=0" "SSDK_OVERLAP_R=37" "SSDK_OVERLAP_R=50" "$OFF" "SSDK_OVERLAP_R=37"; do
  echo "== $cfg" >> gpurun_out/bench_sched.log
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-micro 2>> gpurun_out/bench_sched.err | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: continue
    print(json.dumps({k: j[k] for k in ('value', 'ms_per_step', 'gpu_launches')} | {'e2e': j['e2e']['value'], 'e2e_ms': j['e2e']['ms_per_step'], 'conv_ms': j['roofline']['conv_ms_per_step'], 'frac': j['roofline']['frac'], 'clk': j['clocks']['sm_mhz'], 'why': j['clocks']['reasons']}))
" >> gpurun_out/bench_sched.log
done
cat gpurun_out/e2e_diag.log; tail -12 gpurun_out/pytest_sched.log; cat gpurun_out/bench_sched.log; tail -5 gpurun_out/bench_sched.err
