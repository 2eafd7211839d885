#!/bin/bash
run() { python bench.py --no-cpu-baseline --steps 4 --warmup 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1', 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'K2', round(s['K2 noisemask'],2), 'K3', round(s['K3 tonemask'],2), 'K2+K3', round(s['K2 noisemask']+s['K3 tonemask'],2))"; }
run base
VB200_K2_PERSIST=8 run persist8_noside
for c in 4 5 6 7 8; do VB200_K3_SIDE=1 VB200_K2_PERSIST=$c run side_persist$c; done
