#!/bin/bash
# tools/r2_ab.sh <tag> lib...   : GPU suite once on the default build, then tools/r2_perf.py for the default build and every variant
tag=$1; shift
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$tag.log
echo "== default build"; python tools/r2_perf.py 2>&1 | tee gpurun_out/perf_$tag.log
for so in "$@"; do echo "== $so"; GYMCHESS_B200_LIB=$PWD/$so python tools/r2_perf.py 2>&1 | grep -v "^{" ; done
python tools/list_bench.py
