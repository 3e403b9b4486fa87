#!/bin/bash
# round-2 evidence at HEAD: ncu --set full of the kernels that matter + the launch list of the driver's bench command.
# The profiled batch has its game phases spread (dephase) like the batch bench.py times; launch skip counts: 1 reset + 86
# launches of dephase() + 9 burn-in launches (+ 3 warm-up steps for the single-step kernels).
export GCB_SAMPLED_RANGES=1
NCU="ncu --set full --clock-control none --import-source on -f"
python tools/prof.py --burn-in 576 --steps 64 > gpurun_out/rec_prof_plain.log 2>&1 &&
$NCU -k regex:k_env_step -s 96 -c 1 -o gpurun_out/r2_step_final python tools/prof.py --burn-in 576 --steps 64 > gpurun_out/rec_ncu_step.log 2>&1; echo "step rc=$?"
$NCU -k regex:k_movegen -s 1 -c 1 -o gpurun_out/r2_movegen_final python tools/prof.py --burn-in 576 --steps 64 > gpurun_out/rec_ncu_movegen.log 2>&1; echo "movegen rc=$?"
python tools/prof_single.py 20 > gpurun_out/rec_single_plain.log 2>&1 &&
$NCU -k regex:k_env_step -s 99 -c 1 -o gpurun_out/r2_step1_final python tools/prof_single.py 20 > gpurun_out/rec_ncu_step1.log 2>&1; echo "step1 rc=$?"
python tools/prof_single.py 20 fused > gpurun_out/rec_fused_plain.log 2>&1 &&
$NCU -k regex:k_env_step -s 99 -c 1 -o gpurun_out/r2_step1_mask_final python tools/prof_single.py 20 fused > gpurun_out/rec_ncu_fused.log 2>&1; echo "fused rc=$?"
unset GCB_SAMPLED_RANGES
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/rec_bench_plain.json 2> gpurun_out/rec_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/rec_ncu_launches.log 2>&1; echo "launches rc=$?"
cat gpurun_out/rec_prof_plain.log gpurun_out/rec_single_plain.log gpurun_out/rec_fused_plain.log | grep -v "^{"
