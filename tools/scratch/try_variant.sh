#!/bin/bash
# usage (under gpurun): tools/scratch/try_variant.sh <variant.so> [bench args]   -- runs bench.py with a variant build of the library
V="$1"; shift
cp vorbis_aotuv_lancer_b200/libvorbis_b200.so /tmp/base.so
cp "$V" vorbis_aotuv_lancer_b200/libvorbis_b200.so
python bench.py --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$V', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value'])); print({k: round(v,2) for k,v in d['stage_ms'].items()})"
cp /tmp/base.so vorbis_aotuv_lancer_b200/libvorbis_b200.so
