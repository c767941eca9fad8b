#!/bin/bash
# Profiling pass of one round (run under gpurun, 1 GPU).  Writes only small text files to gpurun_out/:
#   <tag>_launches.csv / .summary.txt   ncu launch list of one replayed step of the bench command
#   <tag>_step_kernels.summary.txt      ncu sections (SOL, memory, launch, occupancy) of the non-conv kernels of a step
#   <tag>_bn_kernels.summary.txt        same for the BatchNorm kernels
#   <tag>_conv_full.summary.txt         ncu --set full of tc_conv_kernel on two north-star layer shapes
# usage: bash tools/profile_round.sh r01b
tag=${1:-prof}
out=gpurun_out
mkdir -p $out
SECT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy"
NONCONV='fps_kernel|modulate|pack_kernel|composite|rep_topk|gather_rows|ball_assign|box_gather|box_scatter|upsample_loss|scatter_rows|trilinear_mix|dilate2|sgemm_kernel|occ_stats|occ_bwd|radix_|lovasz_|label_mode'

timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 950 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
python tools/launch_summary.py $out/${tag}_launches.csv 60 > $out/${tag}_launches.summary.txt 2>&1

timeout 300 ncu $SECT --clock-control none --profile-from-start off -k regex:"$NONCONV" -c 60 \
    -o $out/${tag}_step python tools/ncu_step.py > $out/${tag}_ncu_step.log 2>&1
python tools/ncu_summary.py $out/${tag}_step.ncu-rep > $out/${tag}_step_kernels.summary.txt 2>&1
rm -f $out/${tag}_step.ncu-rep

timeout 200 ncu $SECT --clock-control none --profile-from-start off -k regex:"bn_act" -c 12 \
    -o $out/${tag}_bn python tools/ncu_step.py > $out/${tag}_ncu_bn.log 2>&1
python tools/ncu_summary.py $out/${tag}_bn.ncu-rep > $out/${tag}_bn_kernels.summary.txt 2>&1
rm -f $out/${tag}_bn.ncu-rep

timeout 240 ncu --set full --clock-control none -k regex:tc_conv -s 6 -c 6 -o $out/${tag}_conv \
    python tools/ncu_conv_case.py > $out/${tag}_ncu_conv.log 2>&1
python tools/ncu_summary.py $out/${tag}_conv.ncu-rep > $out/${tag}_conv_full.summary.txt 2>&1
rm -f $out/${tag}_conv.ncu-rep
ls -la $out | head -30
du -sh $out
