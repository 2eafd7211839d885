#!/bin/bash
# usage (under gpurun): tools/prof_round.sh <tag>
# The judged artifacts of a round, all from the default command line of bench.py:
#   bench_<tag>.json            python bench.py                       (our arm, N = 1)
#   bench_reference_<tag>.json  python bench.py --impl reference      (the reference's Lancer build on the host cores)
#   launches_<tag>.csv          ncu launch list (gpu__time_duration.sum) of the same command, 2 steps
#   all_<tag>_raw.csv           ONE ncu --set full capture of every kernel of a whole-workload batch step (64 streams), raw page
#   noisemask_<tag>.ncu-rep     the dominant kernel of that step with its source page
TAG="$1"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# one whole-workload step under --set full: the 29 launches of the warm-up step are skipped, the 29 of the timed step captured
# (the report itself is too large to travel back: its raw page goes home as csv, the report stays on the box)
ncu --set full --clock-control none -k "regex:k_transform|k_noisemask|k_tone|k_mix|k_floor|k_couple|k_residue|k_pack|k_expand|k_stream_scan" \
    -s 29 -c 29 -o /tmp/all_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_all_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_all_$TAG.log
ncu -i /tmp/all_$TAG.ncu-rep --page raw --csv > gpurun_out/all_${TAG}_raw.csv 2>/dev/null
# the dominant kernel (K2, long blocks) once more with its source page, at the same workload
ncu --set full --clock-control none --import-source on -k "regex:k_noisemask" -s 3 -c 1 -o gpurun_out/noisemask_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_noisemask_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_noisemask_$TAG.log
ls -la gpurun_out/ | tail -8
