set -x
G="python tools/bench_batch.py --rows 1000000 --steps 3 --pair"
$G > gpurun_out/r2_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'gemm_filter|rescore|theta' -s 4 -c 2 -o /tmp/prof_r2_gemm -f $G > gpurun_out/r2_ncu_gemm.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2_gemm.ncu-rep gpurun_out/r2_batch_1Mx1024_nq256 --rows 1000000 --dim 1024 --esize 4 > /dev/null 2>&1
G2="python tools/bench_batch.py --rows 1000000 --steps 3 --pair --store mixed"
$G2 > gpurun_out/r2_gemm_bf16_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:'gemm_filter|rescore|theta' -s 4 -c 2 -o /tmp/prof_r2_gemm_bf16 -f $G2 > gpurun_out/r2_ncu_gemm_bf16.log 2>&1
python tools/ncu_summary.py /tmp/prof_r2_gemm_bf16.ncu-rep gpurun_out/r2_batch_bf16_1Mx1024_nq256 --rows 1000000 --dim 1024 --esize 2 > /dev/null 2>&1
