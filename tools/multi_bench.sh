#!/bin/bash
# Multi-GPU bench lines of the BASELINE workloads on one box: tools/multi_bench.sh "<workloads>" "<N list>" [steps]
# writes gpurun_out/mg_<workload>_n<N>.json (one JSON line each)
WLS=${1:-"cfg3 cfg4-split"}; NS=${2:-"1 2 4"}; S=${3:-10}
mkdir -p gpurun_out
for wl in $WLS; do for n in $NS; do
  out=gpurun_out/mg_${wl}_n${n}.json
  if [ $n = 1 ]; then
    python bench.py --gpus 1 --steps $S --warmup 3 --workload $wl --no-cpu-baseline --no-gpu-baseline > $out 2> gpurun_out/mg_${wl}_n${n}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps $S --warmup 3 --workload $wl > $out 2> gpurun_out/mg_${wl}_n${n}.err
  fi
  python - $out <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f frames/s  ms/step %.2f  e2e %.1f  scaling %s  clk %s" % (j["value"], j["ms_per_step"], j["e2e"]["value"], j["scaling"], (j.get("clocks") or {}).get("sm_mhz")))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done; done
