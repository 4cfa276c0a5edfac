nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv,noheader
for v in "default=X=1" "nogate=B200VIT_WINATTN_GATE=0" "oneitem=B200VIT_ATTN_ONE_ITEM=1"; do
  name=${v%%=*}; kv=${v#*=}
  env $kv python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_ms_per_step']; print('$name', round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'), {a: round(b,3) for a,b in k.items() if b>0.2})"
done
