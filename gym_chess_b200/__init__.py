"""gym_chess_b200 -- B200-native (sm_100a) batched drop-in for the gym-chess v2 env step / legal-movegen path.

Public surface:
  BatchedChessEnv   N envs resident in HBM; reset / step / step_index / step_sampled / observe / legal_actions
  PipelinedChessEnv the envs of a device as shards stepped alternately through page-locked 16-bit records (send / recv)
  ChessEngine       drop-in for the reference's PyO3 `ChessEngine` (4 methods, dict/str wire format)
  BatchedChessEngine  the same 4 operations over numpy arrays of positions
  ChessEnvV2        single-env gym-style compat class (reset/step/render/possible_moves/...)
  codec             action <-> move <-> string tables

Everything computes in libgymchess_b200.so (C ABI: include/gymchess_b200.h).  There is no CPU fallback.
"""
from . import codec  # noqa: F401
from ._lib import GcbError, build, lib  # noqa: F401


def __getattr__(name):  # lazy: keeps `import gym_chess_b200` cheap (torch is imported on first use)
    if name in ("BatchedChessEnv",):
        from .batched_env import BatchedChessEnv
        return BatchedChessEnv
    if name == "PipelinedChessEnv":
        from .pipelined import PipelinedChessEnv
        return PipelinedChessEnv
    if name in ("ChessEngine", "BatchedChessEngine"):
        from . import engine
        return getattr(engine, name)
    if name == "ChessEnvV2":
        from .env_v2 import ChessEnvV2
        return ChessEnvV2
    raise AttributeError(name)
