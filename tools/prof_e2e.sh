#!/bin/bash
# usage: tools/prof_e2e.sh <regex> <out-name> [skip] [count]   (like prof.sh but keeps the e2e arm: detector kernels run)
set -e
RX="$1"; OUT="$2"; SKIP="${3:-0}"; CNT="${4:-4}"
CMD="python bench.py --streams 8 --steps 1 --warmup 1 --chunks 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$OUT.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s "$SKIP" -c "$CNT" -o gpurun_out/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
tail -n 2 gpurun_out/ncu_$OUT.log
