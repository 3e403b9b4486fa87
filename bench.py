#!/usr/bin/env python
"""bench.py -- env steps/s (and legal-movegen positions/s) of the batched gym-chess v2 env on B200.

    python bench.py [--gpus N --steps K --warmup W]                 our CUDA path (default N=1)
    python bench.py --impl reference [--gpus N --steps K --warmup W]  the reference algorithm on the host cores
    torchrun ... bench.py --gpus N ...                              one rank per GPU, weak scaling

Workload (BASELINE.json configs[3] "4M envs sharded across 8xB200", per-GPU shard): 524,288 envs per GPU, random
self-play (opponent "none"), uniformly random legal action per ply drawn on the device (Philox4x32-10), auto-reset.
A "step" is one ChessEnvV2.step() of every env of the rank; the fused step kernel runs up to 64 consecutive steps per launch.  `value` = env
steps of all ranks / max-over-ranks device time, state resident in HBM.  `e2e` = the same metric through the
host-buffer C ABI calls (what a binding of the reference env would call): per step the caller's random words cross
PCIe host->device from pinned memory and reward/done/flags device->host, inside the timed region (the kernel reads /
writes the page-locked buffers in place).  The envs of a rank are stepped as two shards with the asynchronous calls
(gcb_env_step_index_packed / gcb_env_wait: 16-bit words in, 16-bit result records out), one in flight while the host
handles the other; the same with the wide int32 / uint8 arrays (e2e.wide_records_value) and the synchronous
one-call-per-step figure (gcb_env_step_index_host, e2e.sync_call_value) are reported next to it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 524288
METRIC = "env_steps_per_sec"
UNIT = "env steps/s"
WORKLOAD = ("random self-play, opponent=none, auto-reset, on-device Philox action draw; %d envs per GPU "
            "(BASELINE.json configs[3]: 4M envs over 8xB200, per-GPU shard)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_baseline(sample_envs, steps_per_env, threads):
    """the oracle (C restatement of lib.rs + chess_v2.py) on the host cores; kind = "port" because the reference's own
    Rust engine cannot be built here (no cargo/rustc)"""
    from oracle import oracle as orc

    t0 = time.time()
    st = orc.selfplay_mt(0, 0, sample_envs, steps_per_env, threads)
    dt = time.time() - t0
    return st["steps"] / dt, dt, st


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # a step = one env step of a bounded sample of the workload's envs
    sample = 8192
    from oracle import oracle as orc

    # the same steady-state workload as the GPU arm: a persistent pool of envs, burn-in so that game phases are mixed
    # (episodes are ~270 plies long), W warm-up steps, then exactly K timed steps of every env of the sample
    pool = orc.SelfplayPool(0, 0, sample)
    pool.run(300, threads)
    pool.run(args.warmup, threads)
    t0 = time.time()
    st = pool.run(args.steps, threads)
    dt = time.time() - t0
    v = st["steps"] / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD % ENVS_PER_GPU, "sample": "%d envs x %d steps after a 300-step burn-in" % (sample, args.steps)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d envs x %d steps of the same self-play workload, C oracle (restates src/lib.rs + "
                                   "chess_v2.py; the Rust engine is unbuildable here), %d pthreads" % (sample, args.steps, threads)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "published_reference": {"steps_per_sec": 3205, "source": "gym_chess/test/v2/test_benchmark.py:46-50 (1851 steps in 0.5776 s, 1 thread, unspecified CPU)"},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--burn-in", type=int, default=600, help="untimed steps so that game phases are mixed")
    ap.add_argument("--no-extras", action="store_true", help="skip the movegen / small-config / cpu_baseline legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(3, args.warmup)  # timing rule: at least 3 warm-up steps (the JSON reports the value used)
    args.steps = max(1, args.steps)
    # stdout carries exactly ONE JSON line: anything a library prints meanwhile (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    from gym_chess_b200 import BatchedChessEnv, _lib

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    N = args.envs_per_gpu
    off, _ = __import__("gym_chess_b200.sharding", fromlist=["x"]).shard_of(rank, world, N)
    env = BatchedChessEnv(N, opponent="none", seed=2, device=local_rank, auto_reset=True, env_id_offset=off)
    L = _lib.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from gym_chess_b200 import sharding

    def max_over_ranks(ms):
        return sharding.max_over_ranks(ms, device=dev)

    # ---- burn-in (game phases mix: episodes are ~270 plies long) + warm-up
    env.step_sampled(args.burn_in)
    env.step_sampled(args.warmup)
    env.reset_stats()
    # ---- timed region 1: kernel-only, state resident in HBM.  Working set (72 B state + 128 B piece slots + 16 KB
    # repetition table per env = 8.7 GB at 524,288 envs) is far larger than the 126 MB L2: no flush needed.
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = L.gcb_launch_count()
    ev0.record()
    env.step_sampled(args.steps)
    ev1.record()
    barrier()
    launches = L.gcb_launch_count() - launches0
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    st = env.stats()
    value = world * N * args.steps / (ms * 1e-3)

    # ---- final NCCL reduce of the episode statistics (the only collective of the job)
    tot = sharding.reduce_stats(env.stats_tensor()).cpu().numpy()

    # ---- timed region 2: end to end through the host-buffer C ABI calls, pinned host memory.  Every env step has its
    # random word copied host->device and its reward / done / flags device->host inside the timed region (the step kernel
    # reads and writes the page-locked buffers in place through PCIe).  The rank's envs are held as TWO shards stepped
    # alternately with the asynchronous calls (gcb_env_step_index_host_async + gcb_env_wait): while the device steps one
    # shard the host has the other's results and issues its next step, so the launch / completion latency of a
    # synchronous call is hidden.  The synchronous single-call figure is reported next to it.
    words = torch.empty((8, N), dtype=torch.int32).pin_memory()
    words.random_(-2 ** 31, 2 ** 31 - 1)
    h_r = torch.empty(N, dtype=torch.int32).pin_memory()
    h_d = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_f = torch.empty(N, dtype=torch.uint8).pin_memory()
    wn, rn, dn, fn = words.numpy().view(np.uint32), h_r.numpy(), h_d.numpy(), h_f.numpy()
    # an episode of random self-play is ~300 steps for 3 envs in 4 (the move cap) and the per-step cost falls as pieces leave
    # the board, so the end-to-end loops cover two whole episode cycles when --steps allows (like the 2000-step kernel region)
    e2e_steps = max(10, min(args.steps, 602))
    for i in range(3):
        env.step_index_host(wn[i % 8], rn, dn, fn)
    barrier()
    ev0.record()
    for i in range(e2e_steps):
        env.step_index_host(wn[i % 8], rn, dn, fn)
    ev1.record()
    barrier()
    sync_ms = max_over_ranks(ev0.elapsed_time(ev1))
    sync_value = world * N * e2e_steps / (sync_ms * 1e-3)

    H = N // 2
    shards = [BatchedChessEnv(H, opponent="none", seed=2, device=local_rank, auto_reset=True, env_id_offset=off + k * H)
              for k in range(2)]
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    import ctypes as C
    vp = C.c_void_p
    ptrs = [[(vp(words[j, k * H:].data_ptr()), vp(h_r[k * H:].data_ptr()), vp(h_d[k * H:].data_ptr()), vp(h_f[k * H:].data_ptr()))
             for j in range(8)] for k in range(2)]
    for k in range(2):
        with torch.cuda.stream(streams[k]):
            shards[k].step_sampled(args.burn_in)

    from gym_chess_b200 import PipelinedChessEnv
    pipe = PipelinedChessEnv(N, shards=2, device=local_rank, env_id_offset=off, opponent="none", seed=2, auto_reset=True)
    pipe.burn_in(args.burn_in)
    words16 = torch.empty((8, 2, H), dtype=torch.int16).pin_memory()   # this rank's inputs of 8 steps, page-locked
    words16.random_(-2 ** 15, 2 ** 15 - 1)
    src16 = [[words16[j, k] for k in range(2)] for j in range(8)]

    def e2e_loop(steps, packed):
        # a step = both shards stepped once (N env steps).  packed: the public pipelined API (PipelinedChessEnv.send_words
        # / recv: uint16 words read from page-locked host memory, uint16 records written back to it; 2 + 2 bytes per env
        # step); else the wide arrays through the asynchronous calls: uint32 words in, int32 reward + uint8 done + uint8
        # flags out (4 + 6 bytes)
        if packed:
            issue = lambda k, i: pipe.send_words(k, src=src16[i % 8][k])
            wait = pipe.recv
        else:
            issue = lambda k, i: shards[k].step_index_host_async(*ptrs[k][i % 8], stream=streams[k])
            wait = lambda k: shards[k].wait(stream=streams[k])
        issue(0, 0)
        for i in range(steps):
            issue(1, i)
            wait(0)   # shard 0's results of step i are in host memory
            if i + 1 < steps:
                issue(0, i + 1)
            wait(1)   # shard 1's results of step i are in host memory

    def timed(packed):
        e2e_loop(3, packed)
        barrier()
        s0, s1 = (pipe.streams[0], pipe.streams[1]) if packed else (streams[0], streams[1])
        ev0.record(s0)
        t0 = time.perf_counter()
        e2e_loop(e2e_steps, packed)
        t1 = time.perf_counter()
        ev1.record(s1)
        barrier()
        # device events (first launch .. last completion) and the host's wall clock around the same loop: the slower one counts
        return max_over_ranks(max(ev0.elapsed_time(ev1), (t1 - t0) * 1e3))

    wide_ms = timed(False)
    e2e_ms = timed(True)
    clk = clocks.stop() if rank == 0 else None  # sampled over the timed regions (kernel-only and end-to-end)
    e2e_value = world * N * e2e_steps / (e2e_ms * 1e-3)
    wide_value = world * N * e2e_steps / (wide_ms * 1e-3)
    e2e_launches = 2 * e2e_steps
    for sh in shards:
        sh.close()
    pipe.close()

    # ---- roofline of the dominant kernel (k_env_step<sampled>): algorithmic bytes per env step, SURVEY.md 8(d):
    # 40 B state read + 40 B state write + 4 B action + 4 B reward + 1 B done + 8 B history append + 8 B x W scanned
    # W = mean repetition window (plies since the last pawn move / capture) the algorithm has to cover; the kernel
    # reads about one hash-table entry per ply instead -- the algorithmic figure stays the survey's.
    W = st["hist_window"] / max(1, st["plies"])
    bytes_per_step = 97.0 + 8.0 * W
    # the step kernel is the only kernel in the timed region; one launch runs up to 64 consecutive steps
    # (gcb_env_step_sampled), so bytes per launch = N * steps-per-launch * bytes_per_step and
    # achieved = bytes per launch / average launch duration = N * bytes_per_step / (ms per step)
    kernel_ms = ms / args.steps
    steps_per_launch = args.steps / max(1, launches)
    achieved = bytes_per_step * N / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = peaks()
    traffic = None  # dram bytes per launch of the step kernel from the committed ncu --set full capture (same N)
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        if tj.get("envs") == N:  # per launch like `achieved`: a launch of the sampled kernel runs up to 64 steps
            traffic = tj["dram_bytes_per_step"] * args.steps / max(1, launches)

    pipes = None  # what actually bounds the kernel (integer pipe), from the committed ncu capture
    pp = os.path.join(ROOT, "profiles", "pipes.json")
    if os.path.exists(pp):
        with open(pp) as f:
            pipes = json.load(f)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD % N, "envs_per_gpu": N, "total_envs": world * N, "burn_in_steps": args.burn_in,
                   "l2": "inputs larger than L2 (8.7 GB resident state per GPU vs 126 MB L2), no flush"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * N, "d2h_bytes_per_step": 2 * N,
                "steps": e2e_steps, "launches": e2e_launches,
                "api": "PipelinedChessEnv.send_words / recv (gcb_env_step_index_packed + gcb_env_wait): uint16 random words "
                       "in, uint16 result records (reward int8 | flags | done) out, page-locked host buffers read / written in "
                       "place by the step kernel; the rank's envs as two shards stepped alternately (one in flight while the "
                       "host handles the other)",
                "wide_records_value": wide_value,
                "wide_records_api": "gcb_env_step_index_host_async: uint32 words in, int32 reward + uint8 done + uint8 flags out "
                                    "(4 + 6 bytes per env step), same two shards",
                "sync_call_value": sync_value,
                "sync_call_api": "gcb_env_step_index_host: one synchronous call per step for all envs of the rank (4 + 6 bytes)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_env_step<MODE_SAMPLED>", "bytes_per_unit": bytes_per_step,
                     "mean_hist_window": W, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": bytes_per_step * N * steps_per_launch, "steps_per_launch": steps_per_launch,
                     "launch_ms": ms / max(1, launches),
                     "note": "integer-pipe (ALU) bound, not HBM bound: see DESIGN.md section 3 and profiles/",
                     "pipes_from_profile": pipes},
        "episode_stats_all_ranks": {k: int(tot[i]) for i, k in enumerate(
            ("steps", "plies", "episodes", "mates", "repetitions", "caps", "wedged", "invalid", "reward_sum", "legal_sum",
             "in_check", "hist_overflow", "slot_overflow", "hist_scanned", "hist_window"))},
    }
    if clk is not None:
        line["clocks"] = clk

    if rank == 0 and world == 1 and not args.no_extras:  # the secondary configs and the CPU baseline: single-GPU runs only
        import ctypes as C
        from gym_chess_b200._lib import Positions, check

        # ---- legal-movegen positions/s on 1,048,576 positions (BASELINE.json configs[1]): the states of 1M envs after
        # a burn-in of random self-play (all game phases, ~6% in check), packed form, kernel-only; output = the
        # reference-ordered uint16 move list + count per position
        M = 1 << 20
        big = BatchedChessEnv(M, opponent="none", seed=7, device=local_rank)
        big.step_sampled(max(200, min(args.burn_in, 400)))
        info = big.info_tensor()
        pl = (info[:, 0] < 0).to(torch.uint8).contiguous()
        rt = (info[:, 1] + 2 * info[:, 2] + 4 * info[:, 3] + 8 * info[:, 4]).to(torch.uint8).contiguous()
        p = big.positions()
        pos = Positions(p.bb01, p.bb23, pl.data_ptr(), rt.data_ptr())
        out = torch.empty((M, 144), dtype=torch.int16, device=dev)
        cnt = torch.empty(M, dtype=torch.int32, device=dev)
        for _ in range(3):
            check(L.gcb_get_possible_moves(M, pos, 0, 0, out.data_ptr(), 144, cnt.data_ptr(), None, None))
        reps = 20
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            check(L.gcb_get_possible_moves(M, pos, 0, 0, out.data_ptr(), 144, cnt.data_ptr(), None, None))
        ev1.record()
        torch.cuda.synchronize()
        mg_ms = ev0.elapsed_time(ev1) / reps
        nl = float(cnt.float().mean().item())
        mg_bytes = 40 + 2 + 2 * nl
        line["movegen"] = {"metric": "legal_movegen_positions_per_sec", "value": M / (mg_ms * 1e-3), "positions": M,
                           "ms_per_launch": mg_ms, "mean_legal": nl, "bytes_per_position": mg_bytes,
                           "hbm_frac": mg_bytes * M / (mg_ms * 1e-3) / 1e9 / peak, "kernel": "k_movegen<false>"}
        big.close()
        del big, out, cnt
        # ---- the same step with the reference-ordered action list of every env materialised after every step
        # (possible_actions decoded from the resident piece slots): what a caller that reads the list each step pays
        lst = torch.empty((N, 144), dtype=torch.int16, device=dev)
        lcnt = torch.empty(N, dtype=torch.int32, device=dev)
        ksteps = max(10, min(args.steps, 100))
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(ksteps):
            env.step_sampled(1)
            check(L.gcb_env_legal_actions(env._h, lst.data_ptr(), 144, lcnt.data_ptr(), None))
        ev1.record()
        torch.cuda.synchronize()
        line["step_plus_action_list"] = {"value": N * ksteps / (ev0.elapsed_time(ev1) * 1e-3), "unit": UNIT,
                                         "note": "k_env_step + k_env_legal_list (uint16[N][144] list + count) per step"}
        # ---- the learner-facing legal-action BIT mask (uint64[N][65], 520 B per env): a pure streaming kernel, the one
        # HBM-bound kernel of the path -- algorithmic bytes = 520 written + 40 state + 8 per own piece (slots) read
        bits = torch.empty((N, 65), dtype=torch.int64, device=dev)
        for _ in range(3):
            check(L.gcb_env_legal_bitmask(env._h, bits.data_ptr(), 65, None))
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(20):
            check(L.gcb_env_legal_bitmask(env._h, bits.data_ptr(), 65, None))
        ev1.record()
        torch.cuda.synchronize()
        bm_ms = ev0.elapsed_time(ev1) / 20
        own_pieces = float(((env.observe().reshape(N, 64) != 0).sum(1).float().mean() / 2).item())
        bm_bytes = 520 + 40 + 8 * own_pieces
        line["legal_bitmask"] = {"kernel": "k_env_legal_bits", "envs": N, "ms_per_launch": bm_ms, "bytes_per_env": bm_bytes,
                                 "achieved_gbs": bm_bytes * N / (bm_ms * 1e-3) / 1e9, "peak_gbs": peak,
                                 "hbm_frac": bm_bytes * N / (bm_ms * 1e-3) / 1e9 / peak,
                                 "note": "bound: hbm; every other kernel of the path is integer-pipe bound"}
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(ksteps):
            env.step_sampled(1)
            check(L.gcb_env_legal_bitmask(env._h, bits.data_ptr(), 65, None))
        ev1.record()
        torch.cuda.synchronize()
        line["step_plus_bitmask"] = {"value": N * ksteps / (ev0.elapsed_time(ev1) * 1e-3), "unit": UNIT,
                                     "note": "k_env_step (one step per launch) + k_env_legal_bits (uint64[N][65] mask) per step"}
        del bits
        # ---- BASELINE.json configs[2]: 65,536 envs on one GPU
        small = BatchedChessEnv(65536, opponent="none", seed=2, device=local_rank)
        small.step_sampled(args.burn_in)
        torch.cuda.synchronize()
        ev0.record()
        small.step_sampled(args.steps)
        ev1.record()
        torch.cuda.synchronize()
        line["config_65536_envs"] = {"value": 65536 * args.steps / (ev0.elapsed_time(ev1) * 1e-3), "unit": UNIT}
        small.close()
        # ---- BASELINE.json configs[4]: repetition / promotion-heavy endgames, 512-slot Zobrist-hash history, 1M envs (the
        # regime in which the repetition windows are long: mean window ~80 plies)
        from gym_chess_b200.boards import endgame_boards
        eg = BatchedChessEnv(1 << 20, opponent="none", seed=5, device=local_rank, initial_boards=endgame_boards(), moves_max=250,
                             history_cap=512)
        eg.step_sampled(600)
        eg.reset_stats()
        torch.cuda.synchronize()
        ev0.record()
        eg.step_sampled(200)
        ev1.record()
        torch.cuda.synchronize()
        egs = eg.stats()
        line["config_endgames_1M_envs"] = {"value": (1 << 20) * 200 / (ev0.elapsed_time(ev1) * 1e-3), "unit": UNIT,
                                           "mean_hist_window": egs["hist_window"] / max(1, egs["plies"]),
                                           "extra_table_probes_per_ply": egs["hist_scanned"] / max(1, egs["plies"]),
                                           "repetitions": egs["repetitions"], "history_cap": 512}
        eg.close()
        # ---- BASELINE.json configs[0]: the reference's own CPU-runnable case -- a single env, random-vs-random self-play,
        # ~1k games -- on ONE host thread (the reference is single-threaded), plus legal-movegen positions/s of the same
        # code on one thread and on all of them.  The engine is the oracle port (the Rust engine cannot be built here).
        from oracle import oracle as orc
        t0 = time.time()
        s1 = orc.selfplay_mt(0, 0, 1, 270000, 1)
        dt1 = time.time() - t0
        sample_b = env.observe()[:65536].reshape(-1, 64).cpu().numpy()
        inf = env.info_tensor()[:65536].cpu().numpy()
        pl, rt = inf[:, 0].astype(np.int8), inf[:, 1:5].astype(np.uint8)
        nthreads = os.cpu_count() or 1
        t0 = time.time()
        orc.movegen_batch(sample_b[:16384], pl[:16384], rt[:16384], False, stride=144, threads=1)
        dtm1 = time.time() - t0
        t0 = time.time()
        orc.movegen_batch(sample_b, pl, rt, False, stride=144, threads=nthreads)
        dtma = time.time() - t0
        line["config_1_cpu"] = {
            "workload": "single env, random self-play (opponent none), 270,000 steps = %d games, 1 thread" % s1["episodes"],
            "env_steps_per_sec_1_thread": s1["steps"] / dt1, "movegen_positions_per_sec_1_thread": 16384 / dtm1,
            "movegen_positions_per_sec_all_threads": 65536 / dtma, "threads": nthreads, "kind": "port",
            "reference_published": {"env_steps_per_sec": 3205, "movegen_positions_per_sec": 4167,
                                    "source": "gym_chess/test/v2/test_benchmark.py:46-50, README.md:372-374 (1 thread, unspecified CPU)"}}
        # ---- CPU baseline on this box's host cores (bounded sample)
        threads = os.cpu_count() or 1
        v, dt, _ = cpu_baseline(128 * threads, 5000, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d envs x 5000 steps of the same self-play workload (%.1f s), C oracle on %d "
                                          "pthreads; reference's published single-thread figure: 3.2e3 steps/s" % (128 * threads, dt, threads)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
