#!/bin/bash
# usage: tools/prof.sh <regex> <out-name> [skip] [count]   (run under gpurun; plain run first, then ncu --set full)
set -e
RX="$1"; OUT="$2"; SKIP="${3:-0}"; CNT="${4:-4}"
CMD="python bench.py --streams 4 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$OUT.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s "$SKIP" -c "$CNT" -o gpurun_out/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
tail -n 2 gpurun_out/ncu_$OUT.log
