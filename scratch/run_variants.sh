#!/bin/bash
# on the GPU box: bench every variant library (kernel table top lines)
cp starch3_b200/libstarch3_b200.so /tmp/orig.so
for v in "$@"; do
  cp scratch/variants/lib$v.so starch3_b200/libstarch3_b200.so
  python bench.py --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/bench_v_$v.log 2>&1
  python - "$v" <<'P'
import json,sys
v=sys.argv[1]
for line in open(f'gpurun_out/bench_v_{v}.log'):
    if line.startswith('{'):
        j=json.loads(line); ks=j['kernels']
        sw=sum(x['ms'] for k,x in ks.items() if 'k_sweep' in k)
        print(v, 'step', round(j['ms_per_step'],2), 'sweeps', round(sw,2), {k:x['ms'] for k,x in ks.items() if any(t in k for t in ('sweep','finish','keys','huff','mtf','zrun'))})
        break
else: print(v,'FAILED'); print(open(f'gpurun_out/bench_v_{v}.log').read()[-600:])
P
done
cp /tmp/orig.so starch3_b200/libstarch3_b200.so
