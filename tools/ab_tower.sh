#!/bin/bash
# In-tower A/B (bench.py cfg2, 30 steps): "name=ENV=VAL[,ENV=VAL]" variants and/or ALT=<tag> library builds.
for i in 1 2; do
for v in "$@"; do
  name=${v%%=*}; kv=${v#*=}; [ "$kv" = "$v" ] && kv="X=1"
  lib=$PWD/rga3-release_b200/libb200vit.so
  [ -f $PWD/rga3-release_b200/libb200vit_$name.so ] && lib=$PWD/rga3-release_b200/libb200vit_$name.so
  env ${kv//,/ } B200VIT_LIB=$lib python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],3), d['kernel_ms_per_step']['attn_full'], d['clocks']['sm_mhz'])"
done; done
