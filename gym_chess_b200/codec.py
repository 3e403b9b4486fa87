"""Action <-> move <-> string codec of the reference (chess_v2.py:492-567, lib.rs:1278-1373) as table lookups."""
import numpy as np

CASTLE_KING_SIDE_WHITE = "CASTLE_KING_SIDE_WHITE"
CASTLE_QUEEN_SIDE_WHITE = "CASTLE_QUEEN_SIDE_WHITE"
CASTLE_KING_SIDE_BLACK = "CASTLE_KING_SIDE_BLACK"
CASTLE_QUEEN_SIDE_BLACK = "CASTLE_QUEEN_SIDE_BLACK"
RESIGN = "RESIGN"
CASTLE_MOVES = [CASTLE_KING_SIDE_WHITE, CASTLE_QUEEN_SIDE_WHITE, CASTLE_KING_SIDE_BLACK, CASTLE_QUEEN_SIDE_BLACK]
NUM_ACTIONS = 64 * 64 + 4 + 1

_SPECIAL = {CASTLE_KING_SIDE_WHITE: 4096, CASTLE_QUEEN_SIDE_WHITE: 4097, CASTLE_KING_SIDE_BLACK: 4098,
            CASTLE_QUEEN_SIDE_BLACK: 4099, RESIGN: 4100}
_SQ = ["abcdefgh"[c] + "87654321"[r] for r in range(8) for c in range(8)]
# table: action -> move string / move tuple
ACTION_TO_STR = [_SQ[a >> 6] + _SQ[a & 63] for a in range(4096)] + CASTLE_MOVES + [RESIGN]
ACTION_TO_MOVE = [((a >> 9, (a >> 6) & 7), ((a >> 3) & 7, a & 7)) for a in range(4096)] + CASTLE_MOVES + [RESIGN]
STR_TO_ACTION = {s: a for a, s in enumerate(ACTION_TO_STR)}
# numpy tables for batch decoding: from-square, to-square (castles/resign: -1)
ACTION_FROM = np.array([a >> 6 for a in range(4096)] + [-1] * 5, np.int8)
ACTION_TO = np.array([a & 63 for a in range(4096)] + [-1] * 5, np.int8)


def move_to_action(move):
    """chess_v2.py:492-506"""
    if type(move) in (list, tuple):
        return (move[0][0] * 8 + move[0][1]) * 64 + (move[1][0] * 8 + move[1][1])
    return _SPECIAL.get(move)


def action_to_move(action):
    """chess_v2.py:508-531"""
    return ACTION_TO_MOVE[int(action)]


def move_to_str_code(move):
    """chess_v2.py:534-540"""
    if move in CASTLE_MOVES:
        return move
    return ACTION_TO_STR[move_to_action(move)]


def str_code_to_move(s):
    """rust_move_to_coords, chess_v2.py:558-567"""
    if s in CASTLE_MOVES:
        return s
    return ACTION_TO_MOVE[STR_TO_ACTION[s]]
