#!/bin/bash
tag=${1:-r2h}
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err ) 2>&1 | grep real; echo "bench rc=$?"
tail -5 gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json")); e = d["e2e"]
print("value %.3e  repeats %s" % (d["value"], ["%.3f" % x for x in d["repeat_ms"]]))
print("e2e %.3e  full_obs %.3e  full_obs_mask %s  sync %.3e" % (e["value"], e["full_obs_value"], e["full_obs_mask_value"], e["sync_call_value"]))
print("roofline", {k: d["roofline"][k] for k in ("frac", "achieved", "bytes_per_unit", "mean_hist_window", "mean_hist_window_steady_state", "steps_per_launch")})
print("launches", d["gpu_launches"], "clocks", d.get("clocks"))
for k in ("other_modes", "movegen", "single_step", "step_plus_action_list", "legal_bitmask", "step_plus_bitmask", "config_65536_envs", "config_endgames_1M_envs", "config_1_cpu", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:700])
r = json.load(open("gpurun_out/bench_ref_$tag.json"))
print("reference arm %.3e  (%s)" % (r["value"], r["cpu_baseline"]["sample"][:120]))
print("ratio e2e/ref %.0f  value/ref %.0f" % (e["value"] / r["value"], d["value"] / r["value"]))
PY
