#!/bin/bash
# round-2 baseline call: GPU suite, driver-style bench, ncu --set full of the four kernels that matter (outputs: gpurun_out/)
tag=${1:-r2a}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
GCB_SAMPLED_RANGES=1 python tools/prof.py --burn-in 576 --steps 64 > gpurun_out/prof_plain_$tag.log 2>&1 &&
GCB_SAMPLED_RANGES=1 ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 10 -c 1 -f -o gpurun_out/r2_step_$tag python tools/prof.py --burn-in 576 --steps 64 > gpurun_out/ncu_step_$tag.log 2>&1
echo "ncu step rc=$?"
python tools/prof_single.py 20 > gpurun_out/prof_single_plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 13 -c 1 -f -o gpurun_out/r2_step1_$tag python tools/prof_single.py 20 > gpurun_out/ncu_step1_$tag.log 2>&1
echo "ncu step1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_movegen -s 1 -c 1 -f -o gpurun_out/r2_movegen_$tag python tools/prof.py --burn-in 576 --steps 64 > gpurun_out/ncu_movegen_$tag.log 2>&1
echo "ncu movegen rc=$?"
python tools/list_bench.py > gpurun_out/list_plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_env_legal_list -s 3 -c 1 -f -o gpurun_out/r2_list_$tag python tools/list_bench.py > gpurun_out/ncu_list_$tag.log 2>&1
echo "ncu list rc=$?"
cat gpurun_out/prof_plain_$tag.log gpurun_out/prof_single_plain_$tag.log gpurun_out/list_plain_$tag.log
