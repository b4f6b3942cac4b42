#!/bin/bash
# Round-2 evidence run on the GPU box (under gpurun): ncu launch list of the bench command, `ncu --set full` of the
# decoder and of the conv3d kernel, conv phase trace.  Summaries are condensed here afterwards (tools/ncu_summary.py).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decoder_tc4 -s 1 -c 1 -f -o gpurun_out/r02_prof_decoder_tc4 \
    env VARIANT=7 python tools/tc_one.py > gpurun_out/r02_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3d_tc -s 11 -c 1 -f -o gpurun_out/r02_prof_conv3d \
    python tools/r02_unet_once.py >> gpurun_out/r02_ncu_full.log 2>&1
VTACO_CV_TRACE=1 python tools/r02_unet_once.py 2>&1 | grep "conv trace" > gpurun_out/r02_conv_trace.txt
tail -2 gpurun_out/r02_ncu_full.log
ls -la gpurun_out | tail -8
