#!/bin/bash
# ncu --set full capture of the decode GEMV kernels (run under gpurun); writes raw CSV pages into gpurun_out/.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemv_bf16_kernel -s 8192 -c 7 -o /tmp/prof_gemv_bf16 -f \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bf16.log 2>&1
ncu -i /tmp/prof_gemv_bf16.ncu-rep --page raw --csv > gpurun_out/prof_gemv_bf16_raw.csv 2>/dev/null
ls -la /tmp/prof_gemv_bf16.ncu-rep
[ $(stat -c %s /tmp/prof_gemv_bf16.ncu-rep) -lt 25000000 ] && cp /tmp/prof_gemv_bf16.ncu-rep gpurun_out/
ncu --set full --clock-control none --import-source on -k regex:"gemv_q_kernel" -s 8192 -c 6 -o /tmp/prof_gemv_q -f \
    python bench.py --workload 1b-w4 --steps 4 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/ncu_q.log 2>&1
ncu -i /tmp/prof_gemv_q.ncu-rep --page raw --csv > gpurun_out/prof_gemv_q_raw.csv 2>/dev/null
[ $(stat -c %s /tmp/prof_gemv_q.ncu-rep) -lt 25000000 ] && cp /tmp/prof_gemv_q.ncu-rep gpurun_out/
python bench.py > gpurun_out/bench_r01_bf16.json 2> gpurun_out/bench_r01_bf16.err
python bench.py --workload 1b-w4 --no-roofline > gpurun_out/bench_r01_w4.json 2> gpurun_out/bench_r01_w4.err
python bench.py --impl reference --steps 32 --warmup 3 > gpurun_out/bench_r01_reference.json 2>/dev/null
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
ls -la gpurun_out
