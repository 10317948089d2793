#!/bin/bash
# Round-2 evidence pass (run under gpurun, 1 GPU): tests, bench lines, launch list, ncu --set full of the dominant kernels.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
(time python -m pytest tests -m gpu -q) > gpurun_out/r02_gputests.log 2>&1; tail -3 gpurun_out/r02_gputests.log
python bench.py > gpurun_out/r02_bench_default.log 2>&1; tail -c 300 gpurun_out/r02_bench_default.log
python bench.py --impl reference > gpurun_out/r02_bench_ref.log 2>&1; tail -c 300 gpurun_out/r02_bench_ref.log
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02_launches_bench_default.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r02_launches_bench_default.log 2>&1
$NCU --set full --import-source on -k regex:scatter_gram_kernel -s 3 -c 1 -o gpurun_out/r02c_fused_c2 -f \
    python tools/prof_fused.py c2 > gpurun_out/r02c_fused_c2.log 2>&1
$NCU --set full --import-source on -k regex:rowpass_bulk -s 4 -c 1 -o gpurun_out/r02b_rowpass_c2 -f \
    python tools/quick_time.py 1000000x100 > gpurun_out/r02b_rowpass_c2.log 2>&1
ls -la gpurun_out | tail -12
