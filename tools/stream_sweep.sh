#!/bin/bash
# sweep the L2 prefetch distance of the streaming decode kernel (tiles per CTA ahead of the ring)
out=gpurun_out/$1; mkdir -p $out
for pf in ${@:2}; do
  MC_STREAM_PF=$pf timeout 200 python bench.py --no-cpu-baseline --no-roofline --steps 256 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('pf=$pf', 'tok/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3))" | tee -a $out/sweep.txt
done
