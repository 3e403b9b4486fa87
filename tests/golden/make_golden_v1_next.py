#!/usr/bin/env python
"""Third pin of the oracle from the reference's own pure-Python env (gym_chess/envs/chess_v1.py, UNMODIFIED): `next_state`.

Run in the BUILD container only (needs /root/reference).  chess_v1.py:366-450 applies a move the same way src/lib.rs:679-784
does -- plain from/to moves with the capture reward (K 0), the "pawn becomes queen" test on the WRONG ends (Q1: only reachable
by a direct call), the four castles as literal square writes -- and v1.py:1008-1026 computes check flags on demand.  This
script records, for positions of random self-play and a few crafted boards: (board, player, move) -> (board after, reward,
white_king_is_checked, black_king_is_checked) for a sample of the legal moves, for synthetic pawn moves onto rows 0 / 7 and
for castle moves whose pieces stand in place.  tests compare the oracle, the host-compiled device rules and the CUDA
engine with these vectors (rights are not compared: v1 tracks them differently).

Output: tests/golden/v1_next_states.json.gz = [{board[64], player, action, board_after[64], reward, checks [w, b] | null}]
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

CASTLES = {4096: "CASTLE_KING_SIDE_WHITE", 4097: "CASTLE_QUEEN_SIDE_WHITE", 4098: "CASTLE_KING_SIDE_BLACK", 4099: "CASTLE_QUEEN_SIDE_BLACK"}


def main():
    _gymshim.install()
    rust = types.ModuleType("gym_chess.gym_chess")
    rust.ChessEngine = orc.OracleEngine  # only so that the reference's __init__.py imports; v1 never calls it
    sys.modules["gym_chess.gym_chess"] = rust
    sys.path.insert(0, "/root/reference")
    from gym_chess.envs.chess_v1 import ChessEnvV1

    rng = np.random.RandomState(17)
    boards, players, _ = ph.harvest_positions(n_envs=30, steps=320, seed=23, every=9)
    keep = rng.permutation(len(boards))[:500]
    cases = [(boards[i].copy(), int(players[i])) for i in keep]
    # castle shapes (pieces in place, both colours) and pawns one step from the "wrong" last rank (Q1)
    for _ in range(60):
        b = np.zeros(64, np.int8)
        b[60], b[56], b[63], b[4], b[0], b[7] = 1, 3, 3, -1, -3, -3
        for s in rng.choice(np.arange(8, 56), size=rng.randint(0, 10), replace=False):
            b[s] = rng.choice([-6, -5, -4, -2, 2, 4, 5, 6])
        cases.append((b, int(rng.choice([1, -1]))))
    out = []
    for b, p in cases:
        if (b == 1).sum() != 1 or (b == -1).sum() != 1:
            continue
        color = "WHITE" if p > 0 else "BLACK"
        try:
            env = ChessEnvV1(opponent="none", log=False, initial_state=b.reshape(8, 8).astype(np.int8).copy())
            legal = env.get_possible_moves(state=env.state, player=color)
        except Exception:  # noqa: BLE001  (v1 raises on adjacent kings)
            continue
        actions = []
        plain = [m for m in legal if not isinstance(m, str)]
        for k in sorted(set([0, len(plain) - 1] + [int(x) for x in rng.randint(0, max(1, len(plain)), size=6)])):
            if 0 <= k < len(plain):
                (r0, c0), (r1, c1) = plain[k]
                actions.append((int(r0) * 8 + int(c0)) * 64 + int(r1) * 8 + int(c1))
        # castles whose rook and king stand in place (the move itself is literal square writes in both implementations)
        if p > 0 and b[60] == 1 and b[63] == 3:
            actions.append(4096)
        if p > 0 and b[60] == 1 and b[56] == 3:
            actions.append(4097)
        if p < 0 and b[4] == -1 and b[7] == -3:
            actions.append(4098)
        if p < 0 and b[4] == -1 and b[0] == -3:
            actions.append(4099)
        # a pawn of the mover pushed "backwards" onto the wrong last rank: the dead promotion branch (Q1)
        own_pawns = np.nonzero(b == 6 * p)[0]
        for s in own_pawns[:2]:
            r0, c0 = divmod(int(s), 8)
            r1 = 7 if p > 0 else 0
            if abs(r1 - r0) == 1:
                actions.append(int(s) * 64 + r1 * 8 + c0)
        for a in actions:
            move = CASTLES[a] if a >= 4096 else ((a >> 9, (a >> 6) & 7), ((a >> 3) & 7, a & 7))
            state = b.reshape(8, 8).astype(np.int8).copy()
            try:
                ns, reward = env.next_state(state, color, move)
            except Exception:  # noqa: BLE001
                continue
            checks = None
            try:
                env.white_king_on_the_board = bool((ns == 1).any())
                env.black_king_on_the_board = bool((ns == -1).any())
                checks = [int(bool(env.king_is_checked(state=ns, player="WHITE"))), int(bool(env.king_is_checked(state=ns, player="BLACK")))]
            except Exception:  # noqa: BLE001  (adjacent kings: v1 raises)
                checks = None
            out.append({"board": [int(x) for x in b], "player": p, "action": int(a), "board_after": [int(x) for x in np.asarray(ns).ravel()],
                        "reward": int(reward), "checks": checks})
    with gzip.GzipFile(os.path.join(HERE, "v1_next_states.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("records", len(out), "with checks", sum(r["checks"] is not None for r in out), "castles", sum(r["action"] >= 4096 for r in out),
          "promotions", sum(r["reward"] >= 10 and r["action"] < 4096 and abs(r["board"][r["action"] >> 6]) == 6 for r in out))
    # agreement with the oracle
    bad = 0
    for r in out:
        b = np.array(r["board"], np.int8)
        ob, orr, oc, rew, st = orc.next_state_batch(b[None], r["player"], np.ones((1, 4), np.uint8), [r["action"]])
        ok = [int(x) for x in ob[0]] == r["board_after"] and int(rew[0]) == r["reward"] and (r["checks"] is None or [int(x) for x in oc[0]] == r["checks"])
        if not ok:
            bad += 1
            if bad <= 5:
                print("DIFF", b.reshape(8, 8), r["player"], r["action"], "\n", np.array(r["board_after"]).reshape(8, 8), ob[0].reshape(8, 8), r["reward"], rew, r["checks"], oc)
    print("disagreements", bad, "of", len(out))


if __name__ == "__main__":
    main()
