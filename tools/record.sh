#!/bin/bash
# One GPU call that refreshes the round's records: tools/record.sh <tag>  (outputs under gpurun_out/)
tag=${1:-vX}
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
GCB_SAMPLED_RANGES=1 ncu --set full --clock-control none --import-source on -k regex:k_env_step -s 10 -c 1 -f -o gpurun_out/r1_step_$tag python tools/prof.py --burn-in 576 --steps 64 > gpurun_out/ncu_step.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_$tag.csv python bench.py --steps 128 --warmup 3 --burn-in 64 --no-extras > gpurun_out/ncu_launches.log 2>&1
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json")); e = d["e2e"]
print(d["value"], e["value"], e["wide_records_value"], e["sync_call_value"], d["roofline"]["frac"])
for k in ("movegen", "legal_bitmask", "step_plus_action_list", "step_plus_bitmask", "config_65536_envs", "config_endgames_1M_envs", "cpu_baseline"):
    print(k, d[k])
print(json.load(open("gpurun_out/bench_ref_$tag.json"))["value"])
PY
