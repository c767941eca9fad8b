#!/bin/bash
# Profiling pass of one round (run under gpurun, 1 GPU).  Writes only small text files to gpurun_out/:
#   <tag>_launches.csv / .summary.txt   ncu launch list of ONE warmed-up eager training step (tools/ncu_step.py:
#                                       fine stage on, FusedAdamW + gradient arena, grad clip)
#   <tag>_step_kernels.summary.txt      ncu sections + DRAM byte counters of the non-conv kernels of that step
#   <tag>_conv_full.summary.txt         ncu --set full of tc_conv_kernel on two north-star layer shapes
#   <tag>_traffic.json                  dram bytes of the dominant kernel (copied to profiles/traffic.json; bench.py
#                                       reports it as roofline.traffic)
# usage: bash tools/profile_round.sh r02
tag=${1:-prof}
out=gpurun_out
mkdir -p $out
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --metrics dram__bytes_read.sum,dram__bytes_write.sum"
NONCONV='fps_kernel|modulate|pack_|composite|rep_topk|gather_rows|ball_assign|box_gather|box_scatter|upsample_loss|scatter_rows|trilinear|dilate2|sgemm_kernel|occ_stats|occ_bwd|radix_scatter|lovasz_apply|label_mode|bn_act|adamw|sample2d|sample3d|gn_rows|select_'

timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/${tag}_launches.csv python tools/ncu_step.py > $out/${tag}_ncu_step.log 2>&1
python tools/launch_summary.py $out/${tag}_launches.csv 80 > $out/${tag}_launches.summary.txt 2>&1

timeout 420 ncu $SECT --clock-control none --profile-from-start off -k regex:"$NONCONV" -c 160 \
    -o $out/${tag}_step python tools/ncu_step.py > $out/${tag}_ncu_step2.log 2>&1
python tools/ncu_summary.py $out/${tag}_step.ncu-rep --top 48 > $out/${tag}_step_kernels.summary.txt 2>&1
rm -f $out/${tag}_step.ncu-rep

timeout 240 ncu --set full --clock-control none -k regex:tc_conv -s 6 -c 6 -o $out/${tag}_conv \
    python tools/ncu_conv_case.py > $out/${tag}_ncu_conv.log 2>&1
python tools/ncu_summary.py $out/${tag}_conv.ncu-rep --traffic-json $out/${tag}_traffic.json > $out/${tag}_conv_full.summary.txt 2>&1
rm -f $out/${tag}_conv.ncu-rep
ls -la $out | tail -12
