#!/bin/bash
# Same-box A/B: round-1 library (_ab/r1) vs the working tree, bench.py resident numbers + per-kernel breakdown.
# usage: tools/ab.sh [steps]   (writes gpurun_out/ab_*.json)
S=${1:-30}
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=j["kernel_ms_per_step"]
    print(sys.argv[1], "ms/step %.3f e2e %.3f clk %s" % (j["ms_per_step"], j["e2e"]["ms_per_step"], j["clocks"]["sm_mhz"]), {a: round(b,3) for a,b in k.items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
if [ -d _ab/r1 ]; then (cd _ab/r1 && python bench.py --no-cpu-baseline --steps $S > ../../gpurun_out/ab_r1.json 2> ../../gpurun_out/ab_r1.err); summ gpurun_out/ab_r1.json; fi
for v in ${VARIANTS:-0 1 2}; do
  B200VIT_RESID_VARIANT=$v python bench.py --no-cpu-baseline --steps $S > gpurun_out/ab_v$v.json 2> gpurun_out/ab_v$v.err; summ gpurun_out/ab_v$v.json
done
if [ -d _ab/r1 ]; then (cd _ab/r1 && python bench.py --no-cpu-baseline --steps $S > ../../gpurun_out/ab_r1b.json 2> ../../gpurun_out/ab_r1b.err); summ gpurun_out/ab_r1b.json; fi
