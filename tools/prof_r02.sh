#!/bin/bash
# Round-2 profiling pass (run under gpurun): launch list of the default bench + ncu --set full of the dominant kernels.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02_launches_bench_default.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r02_launches_bench_default.log 2>&1
$NCU --set full --import-source on -k regex:scatter_gram_kernel -s 3 -c 1 -o gpurun_out/r02_fused_c2 -f \
    python tools/prof_fused.py c2 > gpurun_out/r02_fused_c2.log 2>&1
$NCU --set full --import-source on -k regex:rowpass -s 4 -c 1 -o gpurun_out/r02_rowpass_c2 -f \
    python tools/quick_time.py 1000000x100 > gpurun_out/r02_rowpass_c2.log 2>&1
$NCU --set full --import-source on -k regex:i8_ -s 4 -c 4 -o gpurun_out/r02_i8_k1000 -f \
    python tools/i8_prof.py 262144 1000 > gpurun_out/r02_i8_k1000.log 2>&1
ls -la gpurun_out
