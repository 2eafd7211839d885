#!/bin/bash
# build variants on the GPU box and time them: tools/scratch/exp.sh <file.cu> <stage name> "<flags1>" "<flags2>" ...
F="$1"; ST="$2"; shift 2
cd vorbis_aotuv_lancer_b200/csrc
for v in "$@"; do
  touch $F; make EXTRA="$v" > /dev/null 2>&1
  (cd ../..; python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', '$ST', round(d['stage_ms']['$ST'],3), 'step', round(d['ms_per_step'],2))")
done
