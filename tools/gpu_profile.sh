#!/bin/bash
# Run on the GPU box (under gpurun): smoke, bench, ncu launch list, ncu full capture of the decoder kernels.
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 300 gpurun_out/bench_ref.json
python bench.py --steps 5 --warmup 3 --variant 1 --no-cpu-baseline > gpurun_out/bench_n1_simt.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decoder_tc -s 1 -c 1 -f -o gpurun_out/prof_decoder_tc \
    python tools/tc_one.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
