#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/model_errors.jsonl
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_reference_goldens.py tests/test_gpu_train.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/pytest20.log
for v in "SSDK_BN_AUTO=0" "SSDK_BN_AUTO=1" "SSDK_BN_MAX=128"; do
  tag=$(echo $v | tr '=' '_')
  env $v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-micro > gpurun_out/bench20_$tag.json 2> gpurun_out/bench20_$tag.err
  env $v timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_step20_$tag.csv python tools/profile_step.py step > /dev/null 2>&1
done
tail -3 gpurun_out/pytest20.log; for f in gpurun_out/bench20_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline']['conv_ms_per_step'])"; done
