#!/bin/bash
# weak-scaling records: tools/r2_scale.sh <N> <tag>   (run under gpurun --gpus N)
n=$1; tag=${2:-r2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu_$tag.json 2> gpurun_out/bench_${n}gpu_$tag.err; echo "rc=$?"
tail -3 gpurun_out/bench_${n}gpu_$tag.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 640 --warmup 5 --repeats 3 --no-extras > gpurun_out/bench_${n}gpu_${tag}_long.json 2> gpurun_out/bench_${n}gpu_${tag}_long.err; echo "rc=$?"
python - <<PY
import json
for f in ("gpurun_out/bench_${n}gpu_$tag.json", "gpurun_out/bench_${n}gpu_${tag}_long.json"):
    d = json.load(open(f)); e = d["e2e"]
    print(f, "n_gpus", d["n_gpus"], "value %.3e e2e %.3e full_obs %.3e full_mask %s" % (d["value"], e["value"], e["full_obs_value"], e["full_obs_mask_value"]), d["repeat_ms"])
    print("   modes", {k: "%.3e" % v["value"] for k, v in d["other_modes"].items()}, "episodes", d["episode_stats_all_ranks"]["episodes"])
PY
