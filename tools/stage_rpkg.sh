#!/bin/bash
# Lays the drop-in sources out the way they sit inside the gficf R package:
#   tools/stage_rpkg.sh <gficf-checkout-or-empty-dir>
# copies the replaced / new src/*.cpp, the CUDA sources into src/cuda/ and appends Makevars.cuda to
# src/Makevars (created when absent).  INTEGRATION.md section 1 is the same list in prose.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="${1:?usage: stage_rpkg.sh <package dir>}"
mkdir -p "$DST/src/cuda" "$DST/tests/testthat"
cp "$ROOT"/gficf_b200/rpkg/src/*.cpp "$DST/src/"
cp "$ROOT"/gficf_b200/csrc/gficf_cuda.cu "$ROOT"/gficf_b200/csrc/host_expand.cpp "$ROOT"/gficf_b200/csrc/host_stats.cpp "$ROOT"/gficf_b200/csrc/*.cuh \
   "$ROOT"/gficf_b200/csrc/*.h "$ROOT"/include/gficf_cuda.h "$DST/src/cuda/"
touch "$DST/src/Makevars"
grep -q "cuda/gficf_cuda.o" "$DST/src/Makevars" || cat "$ROOT/gficf_b200/rpkg/src/Makevars.cuda" >> "$DST/src/Makevars"
cp "$ROOT"/gficf_b200/rpkg/tests/testthat/*.R "$DST/tests/testthat/"
echo "staged into $DST"
