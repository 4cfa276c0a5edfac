#!/bin/bash
# MMA-thread / epilogue-warp cycle breakdown (B200_GEMM_TIMING builds) of the residual GEMMs: round 1 vs working tree.
for n in ${NAMES:-proj down}; do
  echo "== r1 $n"; (cd _ab/r1t && python tools/prof_gemm.py $n 2 2>&1 | grep -E "gemm|median" | tail -5)
  for v in ${VARIANTS:-0 1}; do
    echo "== new v$v $n"; B200VIT_LIB=$PWD/rga3-release_b200/libb200vit_timing.so B200VIT_RESID_VARIANT=$v python tools/prof_gemm.py $n 2 2>&1 | grep -E "gemm|median" | tail -5
  done
done
