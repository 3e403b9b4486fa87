#!/bin/bash
# A/B kernel variants on the GPU box: tools/ab.sh <burn-in> lib1.so lib2.so ...   (prints tools/prof.py timings per build)
burn=$1; shift
for so in "$@"; do
  echo "== $so"
  GYMCHESS_B200_LIB=$PWD/$so python tools/prof.py --burn-in $burn --steps 50 2>&1 | grep -E "^(step|movegen)"
done
