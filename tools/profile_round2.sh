#!/bin/bash
# The ncu captures behind profiles/r2_* (one gpurun call on ONE GPU; every ncu run directly after the same command
# exited 0 without ncu).  The reports are summarised ON the box (tools/ncu_summary.py -> gpurun_out/r2_*.md) and only
# the scan's report travels back (gpurun_out/ is limited to 64 MiB).  Usage on the GPU box: bash tools/profile_round2.sh
set -x
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
B="python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline"
$B > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_ncu_list.log 2>&1
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 6 -c 1 -o gpurun_out/prof_r2_scan_10M -f $B > gpurun_out/r2_ncu_scan.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r2_scan_10M.ncu-rep gpurun_out/r2_scan_topk_10Mx1024 --rows 10000000 --dim 1024 --esize 4 \
    --traffic-json gpurun_out/roofline_traffic.json > /dev/null 2>&1
for frac in 0.03 0.07; do
F="python tools/run_scan_once.py --rows 1000000 --dim 1024 --pass-frac $frac --steps 6 --tunable pdl=2"
$F > gpurun_out/r2_filter_${frac}_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'scan_topk|filter_list' -s 6 -c 4 -o /tmp/prof_r2_filter_${frac} -f $F > gpurun_out/r2_ncu_filter_${frac}.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2_filter_${frac}.ncu-rep gpurun_out/r2_filter_list_scan_1Mx1024_${frac} --rows 1000000 --dim 1024 --esize 4 > /dev/null 2>&1
done
G="python tools/bench_batch.py --rows 1000000 --steps 3 --pair"
$G > gpurun_out/r2_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'gemm_filter|rescore|theta' -s 8 -c 4 -o /tmp/prof_r2_gemm -f $G > gpurun_out/r2_ncu_gemm.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2_gemm.ncu-rep gpurun_out/r2_batch_1Mx1024_nq256 --rows 1000000 --dim 1024 --esize 4 > /dev/null 2>&1
G2="python tools/bench_batch.py --rows 1000000 --steps 3 --pair --store mixed"
$G2 > gpurun_out/r2_gemm_bf16_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'gemm_filter|rescore|theta' -s 8 -c 4 -o /tmp/prof_r2_gemm_bf16 -f $G2 > gpurun_out/r2_ncu_gemm_bf16.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2_gemm_bf16.ncu-rep gpurun_out/r2_batch_bf16_1Mx1024_nq256 --rows 1000000 --dim 1024 --esize 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.md
du -sh gpurun_out
