#!/usr/bin/env python
"""Generate tests/golden/render_info.json.gz -- run in the BUILD container only (needs /root/reference).

Replays a selection of the recorded games (trajectories.json.gz) on the reference's UNMODIFIED chess_v2.py shell
(gym shimmed, engine = the C oracle, exactly like make_golden.py) with log=True and records, at reset and after every
replayed step, what the text / info side of the gym surface returns (chess_v2.py:337-353, 422-490, 542-556):
    render(mode="string"), render_moves(possible_moves, mode="string"), [move_to_string(m) for m in possible_moves],
    the whole `info` dict (possible_moves as action codes), and everything the step printed to stdout (log=True:
    the "          >>>>>>>>>> WHITE" header + render_moves([move]) of player_move, chess_v2.py:409-411).
The bot's plies are replayed through a callable opponent (chess_v2.py:171-179), so the BLACK-agent games pin the
callable-opponent path for both colours.

Usage:  python tests/golden/make_golden_render.py
"""
import contextlib
import gzip
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

STEPS_PER_GAME = 60


def info_rec(env, info):
    d = dict(info)
    d["possible_moves"] = [int(env.move_to_action(m)) for m in d["possible_moves"]]
    return {k: (v if isinstance(v, (list, str)) else int(v)) for k, v in d.items()}


def snap(env, info, printed):
    moves = env.possible_moves
    return dict(render=env.render(mode="string"), render_moves=env.render_moves(moves, mode="string"),
                move_strings=[env.move_to_string(m) for m in moves], info=info_rec(env, info), stdout=printed)


def replay(V2, t):
    bots = [t["reset"]["bot_action"]] + [s["bot_action"] for s in t["steps"]]
    cur = {"i": 0}

    def bot(env):
        a = bots[cur["i"]]
        return "resign" if a < 0 else env.action_to_move(a)

    kw = dict(player_color=t["player_color"], opponent=(bot if t["opponent"] == "random" else "none"), log=True,
              initial_board=np.array(t["initial_board"], np.int8).reshape(8, 8))
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        env = V2(**kw)
    rec = dict(reset=snap(env, env.info, out.getvalue()), steps=[])
    for i, s in enumerate(t["steps"][:STEPS_PER_GAME]):
        if s["raised"]:
            break
        cur["i"] = i + 1
        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            state, reward, done, info = env.step(s["action"])
        assert (reward, done) == (s["reward"], s["done"]) and mg.enc_board(env.board) == s["board"], (t["name"], i)
        rec["steps"].append(snap(env, info, out.getvalue()))
    return rec


def main():
    gym_chess = mg.import_reference()
    V2 = gym_chess.envs.chess_v2.ChessEnvV2
    with gzip.open(os.path.join(HERE, "trajectories.json.gz")) as f:
        traj = json.loads(f.read())
    out, seen = {}, set()
    for i, t in enumerate(traj):
        key = (t["name"], t["player_color"], t["opponent"])
        if key in seen:  # one game per (board, mode)
            continue
        seen.add(key)
        out[str(i)] = replay(V2, t)
    mg.dump_gz(out, os.path.join(HERE, "render_info.json.gz"))
    print("render_info: %d games, %d steps" % (len(out), sum(len(r["steps"]) for r in out.values())))


if __name__ == "__main__":
    main()
