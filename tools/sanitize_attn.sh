mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  extra=""; [ $tool = memcheck ] && extra="--leak-check no"
  timeout 500 compute-sanitizer --tool $tool $extra --error-exitcode 99 --print-limit 20 --report-api-errors no \
    python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" -p no:cacheprovider > gpurun_out/san_attn_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/san_attn_$tool.log | tail -3
done
timeout 500 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 99 --print-limit 20 --report-api-errors no \
    python -m pytest tests/test_gpu_tower.py -x -q -m gpu -k "tiny" -p no:cacheprovider > gpurun_out/san_attn_tower.log 2>&1
echo "== memcheck tower tiny rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_attn_tower.log | tail -3
