#!/bin/bash
# usage (under gpurun): [RX=regex SKIP=n CNT=n PICK=k] tools/prof_src.sh <tag> [streams]
# (-k matches the function name without template arguments; a step launches short blocks first, then long: PICK=1)
# Source-level ncu capture of the long-block kernels of one resident batch step; the reports stay on the box
# (too large to travel), their source pages (CUDA + SASS, per-line instruction and stall-sample counts) and raw
# pages come home as csv: gpurun_out/src_<tag>_<kernel>.csv, gpurun_out/raw_<tag>.csv.  Read with tools/ncu_lines.py.
TAG="$1"; STREAMS="${2:-8}"
CMD="python bench.py --streams $STREAMS --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_src_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_src_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on \
    -k "regex:${RX:-k_floor_fit|k_tone_splat|k_tone_fill|k_transform_exact|k_mix|k_floor_encode|k_tone_chase}" \
    -s ${SKIP:-14} -c ${CNT:-14} -o /tmp/src_$TAG $CMD > gpurun_out/ncu_src_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_src_$TAG.log
ncu -i /tmp/src_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_src_$TAG.csv 2>/dev/null
KS="${RX:-k_floor_fit|k_tone_splat|k_tone_fill|k_transform_exact|k_mix|k_floor_encode|k_tone_chase}"
for K in ${KS//|/ }; do
  ncu -i /tmp/src_$TAG.ncu-rep --page source --csv --print-source cuda,sass --kernel-name "regex:$K" --launch-skip ${PICK:-1} --launch-count 1 > gpurun_out/src_${TAG}_$K.csv 2>/dev/null
done
ls -la gpurun_out/ | grep "src_$TAG" | head -20
