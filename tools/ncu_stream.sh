#!/bin/bash
# ncu captures of the streaming decode kernel (one launch = one decode step at KV length 512), bf16 and int4 models:
#   full section set of one launch -> gpurun_out/<dir>/stream_<fmt>.ncu-rep, and the launch list of a short bench run.
out=gpurun_out/${1:-ncu}; mkdir -p $out
for fmt in 0 1; do
  name=$([ $fmt = 0 ] && echo bf16 || echo w4)
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 1 -c 1 -o $out/stream_$name \
      python tools/stream_once.py 1 2 1b 512 $fmt > $out/run_$name.log 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode_stream -c 400 --csv --log-file $out/launches_bench_stream.csv \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_first400.csv \
    python tools/stream_once.py 1 3 1b 16 0 > $out/run_first400.log 2>&1
ls -la $out
