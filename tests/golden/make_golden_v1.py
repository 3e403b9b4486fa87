#!/usr/bin/env python
"""Secondary pin of the oracle: move SETS of the reference's own pure-Python env (gym_chess/envs/chess_v1.py, UNMODIFIED).

Run in the BUILD container only (needs /root/reference).  chess_v1.py is the reference's other implementation of the
same rules; it shares no code with src/lib.rs or with this repo's oracle, so agreement on move sets pins the C
restatement of lib.rs from a second, independent side.  v1 and v2 differ in documented places:
v1 never captures a king with a non-pawn piece, needs both rights for any castle, and raises on adjacent kings.  The
fixture stores v1's answers verbatim; the comparison (tests/test_oracle_golden.py) drops castles and moves onto the
enemy king square from both sides and skips positions v1 refuses.

Output: tests/golden/v1_move_sets.json.gz = [{board[64], player, moves [[from, to], ...], attack [[from, to], ...]}]
(`attack` = v1's get_possible_moves(attack=True): the attack / defence pseudo-moves, in v1's order)
"""
import gzip
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import _gymshim  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from tests import parity_helpers as ph  # noqa: E402


def main():
    _gymshim.install()
    rust = types.ModuleType("gym_chess.gym_chess")
    rust.ChessEngine = orc.OracleEngine  # only so that the reference's __init__.py imports; v1 never calls it
    sys.modules["gym_chess.gym_chess"] = rust
    sys.path.insert(0, "/root/reference")
    from gym_chess.envs.chess_v1 import ChessEnvV1

    boards, players, rights = ph.harvest_positions(n_envs=40, steps=320, seed=11, every=9)
    rng = np.random.RandomState(5)
    keep = rng.permutation(len(boards))[:900]
    out, skipped = [], 0
    for i in keep:
        b, p = boards[i], int(players[i])
        if (b == 1).sum() != 1 or (b == -1).sum() != 1:
            continue
        try:
            env = ChessEnvV1(opponent="none", log=False, initial_state=b.reshape(8, 8).astype(np.int8).copy())
            color = "WHITE" if p > 0 else "BLACK"
            moves = env.get_possible_moves(state=env.state, player=color)
            attack = env.get_possible_moves(state=env.state, player=color, attack=True)
        except Exception:  # noqa: BLE001  (v1 raises on adjacent kings etc.)
            skipped += 1
            continue
        ms = []
        for m in moves:
            if isinstance(m, str):
                ms.append(m)
            else:
                (r0, c0), (r1, c1) = m
                ms.append([int(r0) * 8 + int(c0), int(r1) * 8 + int(c1)])
        att = [[int(r0) * 8 + int(c0), int(r1) * 8 + int(c1)] for (r0, c0), (r1, c1) in attack]
        out.append({"board": [int(x) for x in b], "player": p, "moves": ms, "attack": att})
    with gzip.GzipFile(os.path.join(HERE, "v1_move_sets.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("positions", len(out), "skipped", skipped)
    # report the agreement with the oracle under the documented normalisation
    bad = 0
    for rec in out:
        b = np.array(rec["board"], np.int8)
        exp, cnt = orc.movegen_batch(b[None], np.array([rec["player"]], np.int8), np.zeros((1, 4), np.uint8), False)
        eking = int(np.nonzero(b == -rec["player"])[0][0])
        mine = {(int(a) >> 6, int(a) & 63) for a in exp[0, : cnt[0]] if a < 4096 and (int(a) & 63) != eking}
        theirs = {tuple(m) for m in rec["moves"] if not isinstance(m, str) and m[1] != eking}
        if mine != theirs:
            bad += 1
            if bad <= 5:
                print("DIFF", b.reshape(8, 8), rec["player"], sorted(mine - theirs), sorted(theirs - mine))
    print("disagreements", bad, "of", len(out))


if __name__ == "__main__":
    main()
