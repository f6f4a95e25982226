#!/bin/bash
# Round evidence (run under gpurun): GPU test suite, the bench lines of every workload, launch lists and full-set captures.
out=gpurun_out/${1:-round}
mkdir -p $out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $out/pytest_gpu.log
python bench.py > $out/bench_bf16.json 2> $out/bench_bf16.err
python bench.py --workload 1b-w4 > $out/bench_w4.json 2> $out/bench_w4.err
python bench.py --batch 32 --steps 128 > $out/bench_bf16_batch32.json 2> $out/bench_bf16_batch32.err
python bench.py --workload 1b-bf16-prefill --steps 16 --warmup 3 > $out/bench_prefill_1b.json 2> $out/bench_prefill_1b.err
python bench.py --workload 8b-bf16-prefill --steps 8 --warmup 3 --no-cpu-baseline > $out/bench_prefill_8b.json 2> $out/bench_prefill_8b.err
python bench.py --workload 8b-bf16 --steps 128 --no-cpu-baseline > $out/bench_8b_bf16.json 2> $out/bench_8b_bf16.err
python bench.py --impl reference --steps 16 --warmup 3 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 4450 -c 420 --csv --log-file $out/launches_b32.csv python bench.py --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_b32.log 2>&1
bash tools/ncu_prefill.sh ${1:-round}
ls -la $out
