#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ -- run in the BUILD container only.

What it does (needs /root/reference, which does not exist on the GPU box):
  1. imports the reference's UNMODIFIED gym_chess package (chess_v2.py shell) with
       * `gym` replaced by tests/golden/_gymshim.py (gym is not installed), and
       * `gym_chess.gym_chess.ChessEngine` (the Rust/PyO3 module, unbuildable here: no cargo)
         replaced by oracle.OracleEngine (the C restatement);
  2. runs the reference's own v2 test-suite (gym_chess/test/v2/*.py) against that stack and
     records pass/fail -> reference_v2_tests.json.  Every get_possible_moves /
     get_castle_moves / step call the tests make is logged with its result, so the fixture
     holds positions whose move sets the reference's literal expectations approved;
  3. plays seeded games through the real chess_v2.py `step()` (self-play, WHITE and BLACK agent
     vs a replayable random bot, crafted edge boards, invalid actions, steps after done) and
     records per step (action, bot action, reward, done, board, flags, move_count, legal
     actions) -> trajectories.json.gz.  These pin the oracle's C restatement of
     chess_v2.py:183-294 and, through it, the CUDA step kernel;
  4. harvests positions from those games (both sides, attack on/off) with the move lists the
     reference shell returned -> positions.json.gz.

Usage:  python tests/golden/make_golden.py
"""
import gzip
import importlib
import io
import json
import os
import sys
import contextlib
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import _gymshim  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def import_reference():
    _gymshim.install()
    rust = types.ModuleType("gym_chess.gym_chess")
    rust.ChessEngine = orc.OracleEngine
    sys.modules["gym_chess.gym_chess"] = rust
    sys.path.insert(0, REF)
    import gym_chess  # the reference's real __init__.py

    assert gym_chess.__file__.startswith(REF), gym_chess.__file__
    return gym_chess


def enc_board(board):
    return [int(v) for row in board for v in row]


def flags_of(env):
    return [int(bool(x)) for x in (
        env.white_king_castle_is_possible, env.white_queen_castle_is_possible,
        env.black_king_castle_is_possible, env.black_queen_castle_is_possible,
        env.white_king_is_checked, env.black_king_is_checked)]


# --------------------------------------------------------------------------- 2. the reference's own tests
def run_reference_tests(gym_chess):
    V2 = gym_chess.envs.chess_v2.ChessEnvV2
    calls = []
    orig_gpm, orig_gcm, orig_step = V2.get_possible_moves, V2.get_castle_moves, V2.step

    def state_rec(state):
        return dict(board=enc_board(state["board"]), player=state["current_player"],
                    rights=[int(bool(state[k])) for k in (
                        "white_king_castle_is_possible", "white_queen_castle_is_possible",
                        "black_king_castle_is_possible", "black_queen_castle_is_possible")])

    def gpm(self, state=None, player=None, attack=False):
        st = self.state if state is None else state
        pl = self.current_player if player is None else player
        moves = orig_gpm(self, state=state, player=player, attack=attack)
        calls.append(dict(kind="get_possible_moves", state=state_rec(st), player=pl, attack=bool(attack),
                          actions=[self.move_to_action(m) for m in moves], checks=flags_of(self)[4:]))
        return moves

    def gcm(self, state=None, player=None):
        st = self.state if state is None else state
        pl = self.current_player if player is None else player
        moves = orig_gcm(self, state=state, player=player)
        calls.append(dict(kind="get_castle_moves", state=state_rec(st), player=pl,
                          actions=[self.move_to_action(m) for m in moves]))
        return moves

    def step(self, action):
        before = state_rec(self.state)
        out = orig_step(self, action)
        calls.append(dict(kind="step", state=before, action=int(action), reward=out[1], done=bool(out[2]),
                          board_after=enc_board(self.board), flags_after=flags_of(self)))
        return out

    V2.get_possible_moves, V2.get_castle_moves, V2.step = gpm, gcm, step
    report = []
    names = ["test_basic_moves", "test_capture_moves", "test_king_moves", "test_squares_under_attack",
             "test_castle_moves", "test_run_moves", "test_benchmark"]
    try:
        for modname in names:
            mod = importlib.import_module("gym_chess.test.v2." + modname)
            for fn in sorted(n for n in dir(mod) if n.startswith("test_")):
                first = len(calls)
                ok, err = True, ""
                with contextlib.redirect_stdout(io.StringIO()):
                    try:
                        getattr(mod, fn)()
                    except AssertionError as e:  # only the wall-clock assert of test_benchmark may fail
                        ok, err = False, "AssertionError " + str(e)
                    except Exception as e:  # noqa: BLE001
                        ok, err = False, repr(e)
                rec = dict(module=modname, test=fn, passed=ok, error=err)
                if modname == "test_benchmark":
                    del calls[first:]  # random play, not a golden vector
                else:
                    rec["calls"] = calls[first:]
                report.append(rec)
    finally:
        V2.get_possible_moves, V2.get_castle_moves, V2.step = orig_gpm, orig_gcm, orig_step
    return report


# --------------------------------------------------------------------------- 3. trajectories
class ReplayBot:
    """opponent callable (chess_v2.py:171-179): uniform index into env.possible_moves, logged."""

    def __init__(self, rng):
        self.rng, self.log = rng, []

    def __call__(self, env):
        moves = env.possible_moves
        if len(moves) == 0:
            self.log.append(-1)
            return "resign"  # what make_random_policy returns (chess_v2.py:121-122) -> TypeError downstream
        m = moves[int(self.rng.randint(len(moves)))]
        self.log.append(int(env.move_to_action(m)))
        return m


def B(**pieces):
    """board from {'e1': 1, ...}"""
    b = np.zeros((8, 8), np.int8)
    for sq, v in pieces.items():
        b[8 - int(sq[1]), "abcdefgh".index(sq[0])] = v
    return b


K, Q, R, Bi, N, P = 1, 2, 3, 4, 5, 6

EDGE_BOARDS = {
    "kq_vs_k": B(a8=-K, b5=Q, c6=K),
    "krr_castle": B(e1=K, a1=R, h1=R, c8=-K, h7=-P),
    "rook_a5": B(e1=K, h1=R, a5=R, c8=-K),
    "black_castle_shape": B(e8=-K, a8=-R, h8=-R, e1=K),
    "white_on_rank8": B(e8=K, h8=R, a1=-K),
    "pawn_a7": B(a7=P, e1=K, h8=-K),
    "pawn_jump": B(a2=P, a3=-N, e1=K, e8=-K),
    "pawn_jump_black": B(a7=-P, a6=N, e1=K, e8=-K),
    "king_ray": B(e2=K, e8=-R, a8=-K),
    "kingless_white": B(a8=-K, b2=R, c3=N, h7=-P, g2=P),
    "kingless_both": B(a1=R, h8=-R, b2=P, g7=-P, c3=N, f6=-N),
    "two_white_kings": B(e1=K, c3=K, e8=-K, a8=-R, h4=-Bi),
    "endgame_knights": B(e1=K, e8=-K, b1=N, g8=-N),
    "endgame_bishops": B(e1=K, e8=-K, c1=Bi, c8=-Bi, a2=P, a7=-P),
    "endgame_rooks": B(e1=K, e8=-K, a1=R, h8=-R),
    "queens_many": B(e1=K, e8=-K, a1=Q, b1=Q, c1=Q, a8=-Q, b8=-Q, h5=-Q),
    "pawns_last_rank": B(a8=P, h1=-P, e1=K, e8=-K, b7=P, g2=-P),
    "promo_direct": B(a2=P, e1=K, h8=-K),
}


def play(V2, name, board, player_color, opponent, seed, max_steps, invalid_every=0, after_done=3):
    rng = np.random.RandomState(seed)
    bot = ReplayBot(np.random.RandomState(seed + 7919)) if opponent == "bot" else None
    kw = dict(player_color=player_color, opponent=(bot if bot else "none"), log=False)
    if board is not None:
        kw["initial_board"] = board
    env = V2(**kw)
    rec = dict(name=name, seed=seed, player_color=player_color, opponent=("random" if bot else "none"),
               initial_board=enc_board(env.initial_board), steps=[])
    rec["reset"] = dict(board=enc_board(env.board), flags=flags_of(env), move_count=env.move_count,
                        current_player=env.current_player, legal=[int(a) for a in env.possible_actions],
                        bot_action=(bot.log[-1] if (bot and bot.log) else -1))
    extra = 0
    for t in range(max_steps):
        legal = env.possible_actions
        if invalid_every and t % invalid_every == invalid_every - 1:
            action = int(rng.randint(4101))  # mostly invalid
        elif legal:
            action = int(legal[int(rng.randint(len(legal)))])
        else:
            action = int(rng.randint(4101))
        nbot = len(bot.log) if bot else 0
        raised = False
        try:
            _, reward, done, _ = env.step(action)
        except TypeError:  # bot has no moves (Q9): the reference dies here
            raised = True
            reward, done = None, None
        rec["steps"].append(dict(
            action=action, bot_action=(bot.log[-1] if (bot and len(bot.log) > nbot) else -1),
            reward=reward, done=done, raised=raised, board=enc_board(env.board), flags=flags_of(env),
            move_count=env.move_count, current_player=env.current_player,
            legal=[int(a) for a in env.possible_actions]))
        if raised:
            break
        if done or not env.possible_actions:
            extra += 1
            if extra > after_done:
                break
    return rec


def make_trajectories(gym_chess):
    V2 = gym_chess.envs.chess_v2.ChessEnvV2
    out = []
    for s in range(6):
        out.append(play(V2, "selfplay_default", None, "WHITE", "none", 100 + s, 400))
    for s in range(2):
        out.append(play(V2, "selfplay_default_invalid", None, "WHITE", "none", 200 + s, 200, invalid_every=5))
    for s in range(5):
        out.append(play(V2, "white_vs_bot", None, "WHITE", "bot", 300 + s, 400))
    for s in range(5):
        out.append(play(V2, "black_vs_bot", None, "BLACK", "bot", 400 + s, 500))
    for i, (name, board) in enumerate(EDGE_BOARDS.items()):
        for s in range(2):
            out.append(play(V2, "selfplay_" + name, board, "WHITE", "none", 1000 + 10 * i + s, 400))
        out.append(play(V2, "white_vs_bot_" + name, board, "WHITE", "bot", 2000 + i, 300))
        out.append(play(V2, "black_vs_bot_" + name, board, "BLACK", "bot", 3000 + i, 300))
    # scripted: knight shuffle -> repetition on ply 9
    env = V2(opponent="none", log=False)
    rec = dict(name="knight_shuffle", seed=0, player_color="WHITE", opponent="none",
               initial_board=enc_board(env.initial_board), steps=[])
    rec["reset"] = dict(board=enc_board(env.board), flags=flags_of(env), move_count=0, current_player="WHITE",
                        legal=[int(a) for a in env.possible_actions], bot_action=-1)
    seq = ["g1f3", "g8f6", "f3g1", "f6g8"] * 3
    for m in seq:
        a = orc.str_to_action(m)
        _, reward, done, _ = env.step(a)
        rec["steps"].append(dict(action=a, bot_action=-1, reward=reward, done=done, raised=False,
                                 board=enc_board(env.board), flags=flags_of(env), move_count=env.move_count,
                                 current_player=env.current_player, legal=[int(x) for x in env.possible_actions]))
    out.append(rec)
    return out


# --------------------------------------------------------------------------- 4. positions
def make_positions(gym_chess, trajectories, per_traj=24):
    V2 = gym_chess.envs.chess_v2.ChessEnvV2
    env = V2(opponent="none", log=False)
    rng = np.random.RandomState(5)
    pos, seen = [], set()
    for tr in trajectories:
        steps = tr["steps"]
        if not steps:
            continue
        idx = sorted(set(int(i) for i in rng.randint(0, len(steps), size=min(per_traj, len(steps)))))
        for i in idx:
            st = steps[i]
            key = (tuple(st["board"]), tuple(st["flags"][:4]))
            if key in seen:
                continue
            seen.add(key)
            state = dict(board=np.array(st["board"], np.int8).reshape(8, 8).tolist(), current_player=st["current_player"],
                         white_king_castle_is_possible=bool(st["flags"][0]), white_queen_castle_is_possible=bool(st["flags"][1]),
                         black_king_castle_is_possible=bool(st["flags"][2]), black_queen_castle_is_possible=bool(st["flags"][3]))
            rec = dict(board=st["board"], rights=st["flags"][:4], lists={})
            for player in ("WHITE", "BLACK"):
                for attack in (False, True):
                    moves = env.get_possible_moves(state=state, player=player, attack=attack)
                    rec["lists"]["%s_%d" % (player, int(attack))] = [int(env.move_to_action(m)) for m in moves]
            upd = env.engine.update_state(state)
            rec["update_state"] = [int(bool(upd[k])) for k in (
                "white_king_castle_is_possible", "white_queen_castle_is_possible", "black_king_castle_is_possible",
                "black_queen_castle_is_possible", "white_king_is_checked", "black_king_is_checked")]
            pos.append(rec)
    return pos


def dump_gz(obj, path):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(",", ":")).encode())


def main():
    gym_chess = import_reference()
    report = run_reference_tests(gym_chess)
    with open(os.path.join(HERE, "reference_v2_tests.json"), "w") as f:
        json.dump(report, f, separators=(",", ":"))
    npass = sum(r["passed"] for r in report)
    print("reference v2 tests: %d/%d passed" % (npass, len(report)))
    for r in report:
        if not r["passed"]:
            print("  FAILED", r["module"], r["test"], r["error"])
    traj = make_trajectories(gym_chess)
    dump_gz(traj, os.path.join(HERE, "trajectories.json.gz"))
    print("trajectories:", len(traj), "steps:", sum(len(t["steps"]) for t in traj))
    pos = make_positions(gym_chess, traj)
    dump_gz(pos, os.path.join(HERE, "positions.json.gz"))
    print("positions:", len(pos))


if __name__ == "__main__":
    main()
