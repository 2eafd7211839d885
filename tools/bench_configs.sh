#!/bin/bash
# usage (under gpurun): tools/bench_configs.sh <tag>   -- both arms of bench.py on the other BASELINE configurations
# (C3 q5 stereo, C4 q8 5.1, C5 q10 stereo and 128 kbit/s managed), one JSON line each into gpurun_out/
TAG="$1"
run() {  # name, args...
  local name="$1"; shift
  python bench.py --no-cpu-baseline "$@" > gpurun_out/cfg_${TAG}_${name}.json 2> gpurun_out/cfg_${TAG}_${name}.err || tail -3 gpurun_out/cfg_${TAG}_${name}.err
  python bench.py --impl reference --steps 1 --warmup 0 "$@" > gpurun_out/cfg_${TAG}_${name}_ref.json 2>/dev/null
}
run q5    --config q5_44k_2ch   --streams 64 --seconds 60
run q5_5min --config q5_44k_2ch --streams 24 --seconds 300 --depth 1 --steps 3 --warmup 2
run q8_51 --config q8_48k_6ch   --streams 24 --seconds 60
run q10   --config q10_48k_2ch  --streams 64 --seconds 60
run abr128 --config abr128_44k_2ch --streams 32 --seconds 60 --steps 3 --warmup 2
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/cfg_${TAG}_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("cfg_${TAG}_")[1], round(d["value"]), "audio-s/s", "e2e", round(d["e2e"]["value"]), "kbps", d.get("kbps"))
    except Exception as e:
        print(f, "FAILED", e)
PY
