#!/bin/bash
# Evidence for profiles/: per-kernel ncu --set full captures (summarised ON the box: the reports are too large to travel back),
# launch lists, the N=1 bench line and the CPU reference arm.  Everything lands in gpurun_out/ev/.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/ev
O=gpurun_out/ev
T=/tmp/ncu_tmp; mkdir -p $T
NCU="ncu --clock-control none --import-source on"
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_step.csv python tools/profile_step.py step > $O/profile_step.log 2>&1
python tools/launch_summary.py $O/launches_step.csv > $O/launches_step_summary.txt 2>&1
timeout 900 $NCU --profile-from-start off --set full -k regex:'conv_tcgen05|conv_first' -c 29 -o $T/step_conv_full -f python tools/profile_step.py step > $O/ncu_step_conv.log 2>&1
python tools/ncu_kernel_report.py $T/step_conv_full.ncu-rep > $O/step_conv_per_kernel.md 2>&1
python tools/ncu_summary.py $T/step_conv_full.ncu-rep > $O/step_conv_table.md 2>&1
python tools/ncu_lines.py $T/step_conv_full.ncu-rep conv_first conv 25 conv_first > $O/conv_first_lines.txt 2>&1
timeout 900 $NCU --profile-from-start off --set full -k regex:'nms_kernel|maxpool|topk|dec_prepare|l2norm|preprocess|im2col' -c 14 -o $T/step_small_full -f python tools/profile_step.py step > $O/ncu_step_small.log 2>&1
python tools/ncu_kernel_report.py $T/step_small_full.ncu-rep > $O/step_small_per_kernel.md 2>&1
python tools/ncu_lines.py $T/step_small_full.ncu-rep nms_kernel decode 25 nms_kernelIfLb1 > $O/nms_lines.txt 2>&1
timeout 600 $NCU --profile-from-start off --set full -k regex:'ssd_loss' -c 2 -o $T/loss_full -f python tools/profile_loss.py > $O/ncu_loss.log 2>&1
python tools/ncu_kernel_report.py $T/loss_full.ncu-rep > $O/loss_per_kernel.md 2>&1
python tools/ncu_lines.py $T/loss_full.ncu-rep ssd_loss_kernel loss 25 ssd_loss_kernel > $O/loss_lines.txt 2>&1
timeout 600 $NCU --set full -k regex:'enc_tiles|enc_lb' -c 2 -o $T/enc_micro_full -f python tools/profile_encode.py 256 > $O/ncu_enc_micro.log 2>&1
python tools/ncu_kernel_report.py $T/enc_micro_full.ncu-rep > $O/enc_micro_per_kernel.md 2>&1
python tools/ncu_lines.py $T/enc_micro_full.ncu-rep enc_tiles_kernel encode 30 enc_tiles_kernelILb1 > $O/enc_micro_lines.txt 2>&1
timeout 600 $NCU --set full -k regex:'enc_tiles' -c 1 -o $T/enc_ssd300_full -f python tools/profile_encode300.py > $O/ncu_enc300.log 2>&1
python tools/ncu_kernel_report.py $T/enc_ssd300_full.ncu-rep > $O/enc_ssd300_per_kernel.md 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_micro.csv python tools/profile_step.py micro > $O/profile_micro.log 2>&1
python tools/launch_summary.py $O/launches_micro.csv > $O/launches_micro_summary.txt 2>&1
du -sh $O; ls -la $O | head -40; tail -c 600 $O/bench_n1.json
