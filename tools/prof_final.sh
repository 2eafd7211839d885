#!/bin/bash
# usage (under gpurun): tools/prof_final.sh <tag> [noisemask]
# the judged artifacts of a round: default bench line, reference arm, ncu launch list of the same command and
# (second argument) one --set full capture of the dominant kernel (K2, long blocks) at the bench workload
TAG="$1"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
if [ -n "$2" ]; then
  ncu --set full --clock-control none --import-source on -k "regex:k_noisemask" -s 1 -c 1 -o gpurun_out/noisemask_$TAG python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_noisemask_$TAG.log 2>&1
  tail -n 2 gpurun_out/ncu_noisemask_$TAG.log
fi
ls -la gpurun_out/
