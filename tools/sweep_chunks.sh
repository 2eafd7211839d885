#!/bin/bash
for c in "$@"; do
  python bench.py --no-cpu-baseline --chunks $c --steps 3 2>/dev/null | tail -1 > gpurun_out/chunks_$c.json
  python -c "import json; d=json.load(open('gpurun_out/chunks_$c.json')); print('chunks', $c, 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
