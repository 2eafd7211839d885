#!/bin/bash
for w in "1,2,4,6,8,10,11,11,11" "1,2,4,6,8,8,8,9,9,9" "1,2,3,4,6,8,8,8,8,8,8" "1,2,4,8,10,10,10,8,6,5" "1,2,4,8,12,12,10,8,4,3" "1,2,4,6,8,10,10,8,6,5,4" "1,2,4,6,8,8,8,8,6,5,4,4"; do
  VB200_GROUP_WEIGHTS=$w python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w', 'value', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
