#!/bin/bash
export GCB_SAMPLED_RANGES=1
NCU="ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 13 -c 1 -f"
python tools/prof_single.py 20 > gpurun_out/ps_tile.log 2>&1 && $NCU -o gpurun_out/r2_step1_tile python tools/prof_single.py 20 > gpurun_out/ncu_ps_tile.log 2>&1; echo "tile rc=$?"
GYMCHESS_B200_LIB=$PWD/build_variants/notile.so python tools/prof_single.py 20 > gpurun_out/ps_notile.log 2>&1 && GYMCHESS_B200_LIB=$PWD/build_variants/notile.so $NCU -o gpurun_out/r2_step1_notile python tools/prof_single.py 20 > gpurun_out/ncu_ps_notile.log 2>&1; echo "notile rc=$?"
python tools/prof_single.py 20 fused > gpurun_out/ps_fused.log 2>&1 && $NCU -o gpurun_out/r2_step1_fused python tools/prof_single.py 20 fused > gpurun_out/ncu_ps_fused.log 2>&1; echo "fused rc=$?"
cat gpurun_out/ps_tile.log gpurun_out/ps_notile.log gpurun_out/ps_fused.log
