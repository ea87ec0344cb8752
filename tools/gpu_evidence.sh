#!/bin/bash
# Evidence for profiles/: per-kernel ncu --set full captures, launch lists, the N=1 bench line and the CPU reference arm.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/ev
O=gpurun_out/ev
NCU="ncu --clock-control none --import-source on"
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_step.csv python tools/profile_step.py step > $O/profile_step.log 2>&1
timeout 900 $NCU --profile-from-start off --set full -k regex:'conv_tcgen05|conv_first' -c 29 -o $O/step_conv_full -f python tools/profile_step.py step > $O/ncu_step_conv.log 2>&1
timeout 900 $NCU --profile-from-start off --set full -k regex:'nms_kernel|maxpool|topk|dec_prepare|l2norm|preprocess|im2col' -c 14 -o $O/step_small_full -f python tools/profile_step.py step > $O/ncu_step_small.log 2>&1
timeout 600 $NCU --profile-from-start off --set full -k regex:'ssd_loss' -c 2 -o $O/loss_full -f python tools/profile_loss.py > $O/ncu_loss.log 2>&1
timeout 600 $NCU --set full -k regex:'enc_tiles|enc_lb' -c 2 -o $O/enc_micro_full -f python tools/profile_encode.py 256 > $O/ncu_enc_micro.log 2>&1
timeout 600 $NCU --set full -k regex:'enc_tiles' -c 1 -o $O/enc_ssd300_full -f python tools/profile_encode300.py > $O/ncu_enc300.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_micro.csv python tools/profile_step.py micro > $O/profile_micro.log 2>&1
ls -la $O | head -30; tail -c 600 $O/bench_n1.json; cat $O/bench_reference.json | cut -c1-600
