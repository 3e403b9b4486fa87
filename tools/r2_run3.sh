#!/bin/bash
tag=${1:-r2c}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
python tools/r2_perf.py > gpurun_out/perf_$tag.log 2>&1; echo "perf rc=$?"; cat gpurun_out/perf_$tag.log
for c in 2 4 16 32; do echo "== GCB_RUN_CHUNK=$c"; GCB_RUN_CHUNK=$c python tools/r2_perf.py 2>&1 | grep -E "sampled run|65,536"; done
