#!/bin/bash
# The ncu captures behind profiles/r1b_* (one gpurun call; every ncu run directly after the same command exited 0
# without ncu).  Usage on the GPU box: bash tools/profile_round.sh [quick]
set -x
B="python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/r1b_plain.json 2> gpurun_out/r1b_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_bench.csv $B > gpurun_out/r1b_ncu_list.log 2>&1
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 4 -c 2 -o gpurun_out/prof_r1b_scan_10M -f $B > gpurun_out/r1b_ncu_scan.log 2>&1
[ "$1" = quick ] && exit 0
F="python tools/run_scan_once.py --rows 10000000 --dim 1024 --pass-frac 0.07 --steps 6"
$F > gpurun_out/r1b_filter_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'scan_topk|filter_list' -s 6 -c 4 -o gpurun_out/prof_r1b_filter7 -f $F > gpurun_out/r1b_ncu_filter.log 2>&1
G="python tools/bench_batch.py --rows 1000000 --steps 3"
$G > gpurun_out/r1b_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_filter|rescore|theta' -s 8 -c 4 -o gpurun_out/prof_r1b_gemm -f $G > gpurun_out/r1b_ncu_gemm.log 2>&1
G2="python tools/bench_batch.py --rows 1000000 --steps 3 --store mixed"
$G2 > gpurun_out/r1b_gemm_bf16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_filter|rescore|theta' -s 8 -c 4 -o gpurun_out/prof_r1b_gemm_bf16 -f $G2 > gpurun_out/r1b_ncu_gemm_bf16.log 2>&1
H="python tools/run_scan_once.py --rows 12500000 --dim 768 --store bf16 --steps 6"
$H > gpurun_out/r1b_bf16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 3 -c 2 -o gpurun_out/prof_r1b_scan_bf16 -f $H > gpurun_out/r1b_ncu_bf16.log 2>&1
ls -la gpurun_out/*.ncu-rep
