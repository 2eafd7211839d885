#!/bin/bash
# usage (under gpurun): tools/prof_all.sh <out-name> [streams]
# plain bench first, then one ncu --set full capture of every kernel launch of one batch step at a reduced stream count
OUT="$1"; NS="${2:-16}"
CMD="python bench.py --streams $NS --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_$OUT.log 2>&1 || { tail -5 gpurun_out/plain_$OUT.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:k_transform|k_noisemask|k_tone|k_mix|k_floor|k_couple|k_residue|k_pack" -s 30 -c 30 -o gpurun_out/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
tail -n 2 gpurun_out/ncu_$OUT.log
