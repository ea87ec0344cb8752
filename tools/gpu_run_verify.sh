#!/bin/bash
# End-of-round verification + evidence of the default build: full GPU test suite, smoke, the N=1 bench line with micro rows and the
# CPU arm, the reference arm, the ncu launch list of one step and a --set full capture of its conv launches (summarised on the box).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/ev
O=gpurun_out/ev
T=/tmp/ncu_tmp; mkdir -p $T
rm -f gpurun_out/model_errors.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -40 > $O/pytest_verify.log
echo "exit ${PIPESTATUS[0]}" >> $O/pytest_verify.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_verify.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_step.csv python tools/profile_step.py step > $O/profile_step.log 2>&1
python tools/launch_summary.py $O/launches_step.csv > $O/launches_step_summary.txt 2>&1
timeout 900 ncu --clock-control none --import-source on --profile-from-start off --set full -k regex:'conv_tcgen05|conv_first' -c 29 -o $T/step_conv_full -f python tools/profile_step.py step > $O/ncu_step_conv.log 2>&1
python tools/ncu_kernel_report.py $T/step_conv_full.ncu-rep > $O/step_conv_per_kernel.md 2>&1
python tools/ncu_summary.py $T/step_conv_full.ncu-rep > $O/step_conv_table.md 2>&1
timeout 240 python tools/enc_bench.py 256 > $O/enc_bench.log 2>&1
tail -6 $O/pytest_verify.log; tail -2 $O/smoke_verify.log; tail -c 1500 $O/bench_n1.json; echo; tail -c 400 $O/bench_reference.json; echo; head -20 $O/launches_step_summary.txt; tail -3 $O/ncu_step_conv.log
