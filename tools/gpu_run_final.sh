cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/ev2; O=gpurun_out/ev2
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_train_dist.py 2>&1 | tail -6 > $O/pytest_verify.log
echo "exit ${PIPESTATUS[0]}" >> $O/pytest_verify.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_verify.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_step.csv python tools/profile_step.py step > $O/profile_step.log 2>&1
python tools/launch_summary.py $O/launches_step.csv > $O/launches_step_summary.txt 2>&1
tail -3 $O/pytest_verify.log; tail -1 $O/smoke_verify.log; head -c 600 $O/bench_n1.json; echo; head -14 $O/launches_step_summary.txt
