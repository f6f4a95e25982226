#!/bin/bash
# ncu captures of the tensor-core prompt path (run under gpurun): the launch list of one prompt step and one full-set capture
# each of the tcgen05 GEMM (w13 shape inside the step) and of the causal attention kernel; raw CSV pages go to gpurun_out/$1.
out=gpurun_out/${1:-ncu_prefill}
mkdir -p $out
BENCH="python bench.py --workload 1b-bf16-prefill --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_prefill.csv $BENCH > $out/ncu_launches.log 2>&1
# skip the first prompt (130 launches): kernels of the second one
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 68 -c 4 -o /tmp/prof_gemm_tc -f $BENCH > $out/ncu_gemm.log 2>&1
ncu -i /tmp/prof_gemm_tc.ncu-rep --page raw --csv > $out/prof_gemm_tc_raw.csv 2>/dev/null
[ $(stat -c %s /tmp/prof_gemm_tc.ncu-rep) -lt 25000000 ] && cp /tmp/prof_gemm_tc.ncu-rep $out/
timeout 400 ncu --set full --clock-control none --import-source on -k regex:prefill_attn_kernel -s 20 -c 1 -o /tmp/prof_pattn -f $BENCH > $out/ncu_attn.log 2>&1
ncu -i /tmp/prof_pattn.ncu-rep --page raw --csv > $out/prof_pattn_raw.csv 2>/dev/null
[ $(stat -c %s /tmp/prof_pattn.ncu-rep) -lt 25000000 ] && cp /tmp/prof_pattn.ncu-rep $out/
ls -la $out
