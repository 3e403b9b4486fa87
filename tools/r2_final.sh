#!/bin/bash
# final records of the round on one GPU: GPU suite, smoke, reference arm, bench (driver's command line), bench long run
tag=${1:-r2final}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_1gpu_$tag.json 2> gpurun_out/bench_1gpu_$tag.err; echo "bench rc=$?"
python bench.py --gpus 1 --steps 640 --warmup 5 --repeats 3 --no-extras > gpurun_out/bench_1gpu_${tag}_long.json 2> gpurun_out/bench_1gpu_${tag}_long.err; echo "bench long rc=$?"
python - <<PY
import json
for f in ("gpurun_out/bench_1gpu_$tag.json", "gpurun_out/bench_1gpu_${tag}_long.json"):
    d = json.load(open(f)); e = d["e2e"]
    print("value %.3e  repeats %s" % (d["value"], ["%.3f" % x for x in d["repeat_ms"]]))
    print("e2e %.3e  full_obs %.3e  full_obs_mask %s  sync %.3e  frac %.3f" % (e["value"], e["full_obs_value"], e["full_obs_mask_value"], e["sync_call_value"], d["roofline"]["frac"]))
    print("modes", {k: "%.3e" % v["value"] for k, v in d["other_modes"].items()})
    for k in ("movegen", "single_step", "step_plus_action_list", "legal_bitmask", "step_plus_bitmask", "config_65536_envs", "config_endgames_1M_envs", "cpu_baseline"):
        if k in d: print(k, json.dumps(d.get(k))[:420])
r = json.load(open("gpurun_out/bench_ref_$tag.json"))
print("reference arm %.3e" % r["value"])
PY
