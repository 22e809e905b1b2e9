#!/bin/bash
# usage: bench_brief.sh [steps] [extra bench.py flags]  -- prints the few bench fields that matter when iterating on a kernel
steps=${1:-100}; shift
python bench.py --steps $steps --warmup 5 --no-cpu-baseline "$@" 2>/tmp/bench_brief.err | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t); r=d['roofline']
    print({'fps':d['value'],'ms':d['ms_per_step'],'e2e':d['e2e']['value'],'stencil_us':r['avg_launch_us'],'frac':r['frac'],'GBs':r['achieved'],'E':r['E_per_launch'],'U':r['U_per_launch'],'launches':d['gpu_launches']})
except Exception as e:
    print('bench failed', e, t[-300:]); print(open('/tmp/bench_brief.err').read()[-1500:])"
