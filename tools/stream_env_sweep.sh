#!/bin/bash
# usage: stream_env_sweep.sh <outdir> <ENVVAR> v1 v2 ... : bench the streaming kernel under different values of one knob
out=gpurun_out/$1; mkdir -p $out; var=$2
for v in ${@:3}; do
  env $var=$v timeout 200 python bench.py --no-cpu-baseline --no-roofline --steps 256 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$var=$v', 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))" | tee -a $out/sweep_$var.txt
done
