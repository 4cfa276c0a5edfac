#!/bin/bash
# A/B of the full-attention kernel: kernel durations (ncu, no replay metrics) for S = 1024 and S = 2304, for each
# "name=ENV=VALUE" variant given on the command line (default: with and without the MUFU token), optionally against another
# library build (ALT=<tag> -> libb200vit_<tag>.so).
mkdir -p gpurun_out
VARIANTS=${@:-"token=B200VIT_ATTN_TOKEN=1 notoken=B200VIT_ATTN_TOKEN=0"}
for s in 1024 2304; do
  for v in $VARIANTS; do
    name=${v%%=*}; kv=${v#*=}
    env ${kv//,/ } ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ -s 3 -c 6 --csv --log-file gpurun_out/attn_ab.csv python tools/prof_attn.py $s 2 > /dev/null 2>&1
    echo "seg $s $name: $(grep attn_ gpurun_out/attn_ab.csv | awk -F'","' '{gsub(/"/,"",$NF); printf "%s ", $NF}')"
  done
  if [ -n "$ALT" ]; then
    B200VIT_LIB=$PWD/rga3-release_b200/libb200vit_$ALT.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ -s 3 -c 6 --csv --log-file gpurun_out/attn_ab.csv python tools/prof_attn.py $s 2 > /dev/null 2>&1
    echo "seg $s $ALT: $(grep attn_ gpurun_out/attn_ab.csv | awk -F'","' '{gsub(/"/,"",$NF); printf "%s ", $NF}')"
  fi
done
for tag in $EXTRA_LIBS; do
  for s in 1024; do
    B200VIT_LIB=$PWD/rga3-release_b200/libb200vit_$tag.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ -s 3 -c 6 --csv --log-file gpurun_out/attn_ab.csv python tools/prof_attn.py $s 2 > /dev/null 2>&1
    echo "seg $s $tag: $(grep attn_ gpurun_out/attn_ab.csv | awk -F'","' '{gsub(/"/,"",$NF); printf "%s ", $NF}')"
  done
done
