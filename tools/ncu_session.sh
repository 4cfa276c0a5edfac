#!/bin/bash
# One profiling session on the B200 (tools/ncu_session.sh [tag]): the ncu launch list of two steady-state bench steps,
# and an `ncu --set full` capture of every kernel that matters, each reduced on the box to the handful of metrics the
# roofline discussion uses.  Raw .ncu-rep files stay in gpurun_out/ (scratch); the summaries go to profiles/.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
KREGEX='regex:gemm_tcgen05|attn_tc|attn_full_kernel|rmsnorm_kernel|overlay_patchify|cast_bf16|raster_kernel'
if [ -z "$NO_LAUNCH_LIST" ]; then
# ---- launch list: skip the warm-up forwards (138 launches each incl. none from torch), take two forwards
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 552 -c 276 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/${TAG}_launches.csv)"
fi
METRICS='gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |dram__bytes_read.sum,|dram__bytes_write.sum,|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|sm__throughput.avg.pct|lts__t_bytes.sum |lts__t_bytes.sum,|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum |l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,|smsp__cycles_active.avg|sm__cycles_elapsed.max|smsp__inst_executed.sum |smsp__inst_executed.sum,|lts__t_sector_hit_rate.pct|sm__pipe_tensor_op_hmma|smsp__issue_active.avg.pct'
cap() {   # name, kernel regex, skip, command...
  local name=$1 kre=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k "regex:$kre" -s $skip -c 1 -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i $OUT/${TAG}_$name.ncu-rep --page raw --csv 2>/dev/null > $OUT/${TAG}_${name}_raw.csv
  python - "$OUT/${TAG}_${name}_raw.csv" "$OUT/${TAG}_ncu_${name}.csv" <<'PY'
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
if len(rows) < 3:
    print("  no data"); sys.exit(0)
hdr, units, vals = rows[0], rows[1], rows[2]
keep = re.compile(r"Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__pipe_tensor_cycles_active.avg.pct|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|sm__throughput.avg.pct_of_peak_sustained_elapsed|lts__t_bytes.sum$|lts__t_sector_hit_rate.pct|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|sm__cycles_elapsed.max|smsp__inst_executed.sum$|smsp__issue_active.avg.pct_of_peak_sustained_active|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__average_warp.*stall|launch__shared_mem_per_block|sm__pipe_tensor_op|dram__throughput|lts__throughput.avg.pct")
with open(sys.argv[2], "w") as f:
    f.write("metric,unit,value\n")
    for h, u, v in zip(hdr, units, vals):
        if keep.search(h):
            f.write(f"{h},{u},{v}\n")
print("  ->", sys.argv[2])
PY
}
CAPS=${CAPS:-"gateup qkv_winattn qkv proj down attn_full overlay"}   # which --set full captures to take
for c in $CAPS; do
  case $c in
    gateup)      cap gateup 'gemm_tcgen05' 8 python tools/prof_gemm.py gateup 2 ;;
    qkv_winattn) cap qkv_winattn 'gemm_tcgen05' 8 python tools/prof_gemm.py qkvwin 2 ;;
    qkv)         cap qkv 'gemm_tcgen05' 8 python tools/prof_gemm.py qkv 2 ;;
    proj)        cap proj 'gemm_tcgen05' 8 python tools/prof_gemm.py proj 2 ;;
    down)        cap down 'gemm_tcgen05' 8 python tools/prof_gemm.py down 2 ;;
    attn_full)   cap attn_full 'attn_full_kernel' 4 python tools/prof_attn.py 1024 2 ;;
    overlay)     cap overlay 'overlay_patchify' 6 python tools/prof_overlay.py ;;
  esac
done
ls -la $OUT | grep ${TAG}_ | head -40
