#!/bin/bash
# usage: tools/sweep_env.sh VAR v1 v2 ...   (bench stage times under different values of an experiment variable)
VAR="$1"; shift
for v in "$@"; do
  env $VAR=$v python bench.py --no-cpu-baseline --steps 3 2>/dev/null | tail -n 1 > gpurun_out/env_$v.json
  python -c "import json; d=json.load(open('gpurun_out/env_$v.json')); print('$VAR', '$v', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), {k: round(x,2) for k,x in d['stage_ms'].items()})"
done
