#!/bin/bash
# usage (under gpurun): tools/prof_round.sh <tag>
# The judged artifacts of a round, all from the default command line of bench.py:
#   bench_<tag>.json            python bench.py                       (our arm, N = 1)
#   bench_reference_<tag>.json  python bench.py --impl reference      (the reference's Lancer build on the host cores)
#   launches_<tag>.csv          ncu launch list (gpu__time_duration.sum) of the same command, 2 steps
#   all_<tag>.ncu-rep           ONE ncu --set full capture of every kernel of a whole-workload batch step (64 streams)
TAG="$1"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# one whole-workload step under --set full: skip the segmenter's launches of the set-up phase and the warm-up step
ncu --set full --clock-control none --import-source on -k "regex:k_transform|k_noisemask|k_tone|k_mix|k_floor|k_couple|k_residue|k_pack|k_expand|k_stream_scan" \
    -s 40 -c 40 -o gpurun_out/all_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_all_$TAG.log 2>&1
tail -n 2 gpurun_out/ncu_all_$TAG.log
ls -la gpurun_out/ | tail -8
