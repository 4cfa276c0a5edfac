#!/bin/bash
# compute-sanitizer pass over the kernels: memcheck, racecheck, synccheck on the per-kernel parity
# tests and the tiny whole-path tests.  Writes gpurun_out/sanitizer_<tool>.log; summary lines are copied to profiles/.
mkdir -p gpurun_out
SEL='gemm or attention or rmsnorm or overlay or patchify or cast or raster or polygons or lines'
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  echo "== $tool ops"
  timeout ${LIMIT:-900} compute-sanitizer --tool $tool $extra --error-exitcode 99 --print-limit 20 --report-api-errors no \
    python -m pytest tests/test_gpu_ops.py tests/test_gpu_raster.py -x -q -m gpu -k "$SEL" -p no:cacheprovider \
    > gpurun_out/sanitizer_${tool}_ops.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_ops.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitizer_${tool}_ops.log | tail -4
  echo "== $tool tower"
  timeout ${LIMIT:-900} compute-sanitizer --tool $tool $extra --error-exitcode 99 --print-limit 20 --report-api-errors no \
    python -m pytest tests/test_gpu_tower.py -x -q -m gpu -k "tiny or golden or odd_frame or splice" -p no:cacheprovider \
    > gpurun_out/sanitizer_${tool}_tower.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_tower.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" gpurun_out/sanitizer_${tool}_tower.log | tail -4
done
