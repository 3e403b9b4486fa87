"""Exercise every env kernel on awkward env counts with guard regions around every env array (GCB_GUARD_BYTES) and --
when GYMCHESS_B200_LIB points at the CHECKED build -- index checks compiled into the kernels.  Prints one JSON line:
{"checked": 0|1, "violations": bit set of failed index checks, "guard_bad_bytes": n, "envs": [...]}.
Run by tests/test_memory_safety.py (in-process for the product build, as a subprocess for the checked builds)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def exercise(sizes, quick=False):
    os.environ["GCB_GUARD_BYTES"] = "4096"
    import numpy as np
    import torch

    from gym_chess_b200 import BatchedChessEnv, _lib
    from gym_chess_b200.boards import endgame_boards

    L = _lib.lib()
    bad_total = 0
    rng = np.random.RandomState(1)
    for N in sizes:
        for kw in (dict(opponent="none"), dict(opponent="random"), dict(opponent="random", player_color="BLACK"),
                   dict(opponent="external", player_color="BLACK"),
                   dict(opponent="none", initial_boards=endgame_boards(), history_cap=8, moves_max=250)):
            if quick and kw.get("opponent") != "none":
                continue
            env = BatchedChessEnv(N, seed=3, **kw)
            if kw["opponent"] == "external":
                env.bot_ply(torch.full((N,), 3364, dtype=torch.int32))          # e2e4 for every env
                legal, cnt = env.legal_actions()
                env.step(legal[:, 0].to(torch.int32) & 0xFFFF)
                env.bot_ply(torch.full((N,), 3299, dtype=torch.int32))          # d2d4
            else:
                env.step_sampled(1)
                env.step_sampled(70)                                             # two launches, multi-step kernel
                legal, cnt = env.legal_actions()
                acts = (legal[:, 0].to(torch.int32) & 0xFFFF)
                acts[::7] = 4100                                                 # some invalid actions
                env.step(acts)
                env.step_index(torch.from_numpy(rng.randint(0, 2 ** 31, size=N).astype(np.int32)).cuda())
                env.step_index_host(rng.randint(0, 2 ** 32, size=N, dtype=np.uint64).astype(np.uint32))   # pageable: staged chunks
                pin16 = torch.zeros(N, dtype=torch.int16).pin_memory()
                out16 = torch.zeros(N, dtype=torch.int16).pin_memory()
                env.step_index_packed(pin16, out16)
                env.wait()
            env.observe(), env.info_tensor(), env.legal_bitmask(), env.legal_actions()
            if N <= 70001:
                env.legal_mask()
            m = (torch.arange(N, device="cuda") % 3 == 0).to(torch.uint8)
            env.reset(m)
            b = env.observe().reshape(N, 64)
            info = env.info_tensor()
            env.set_state(b, info[:, 0].to(torch.int8), info[:, 1:5].to(torch.uint8), None, m)
            snap = env.snapshot()
            env.step_sampled(3) if kw["opponent"] != "external" else None
            env.restore(snap)
            env.stats()
            torch.cuda.synchronize()
            nbad = C.c_uint64()
            _lib.check(L.gcb_env_check_guards(env._h, C.byref(nbad)))
            if nbad.value:
                print("guard damage N=%d %s: %s" % (N, kw, L.gcb_last_error().decode()), file=sys.stderr)
            bad_total += nbad.value
            env.close()
    v = C.c_uint64()
    _lib.check(L.gcb_debug_violations(C.byref(v), 1))
    return dict(checked=int(L.gcb_build_is_checked()), violations=int(v.value), guard_bad_bytes=int(bad_total), envs=list(sizes))


if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [1, 33, 64, 160, 3000]
    print(json.dumps(exercise(sizes, quick=len(sys.argv) > 2)))
