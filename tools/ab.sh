#!/bin/bash
# Same-box A/B: the round-1 library vs the working tree, bench.py resident numbers + per-kernel breakdown.
# usage: tools/ab.sh [steps]   (writes gpurun_out/ab_*.json)
# The round-1 build lives in _ab/r1 (git-ignored, travels with gpurun):
#   git worktree add _ab/r1 8246226 && (cd _ab/r1 && python rga3-release_b200/build.py)
# For A/B of env switches or tagged builds of THIS tree use tools/ab_tower.sh.
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
run_r1() { if [ -d _ab/r1 ]; then (cd _ab/r1 && python bench.py --no-cpu-baseline --steps $S > ../../gpurun_out/ab_r1$1.json 2> ../../gpurun_out/ab_r1$1.err); summ gpurun_out/ab_r1$1.json; fi; }
run_r1 ""
python bench.py --no-cpu-baseline --no-gpu-baseline --steps $S > gpurun_out/ab_now.json 2> gpurun_out/ab_now.err; summ gpurun_out/ab_now.json
run_r1 b
