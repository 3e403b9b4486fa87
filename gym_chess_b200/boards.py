"""Initial-board presets (reference wire format: int8[64], row 0 = rank 8; K1 Q2 R3 B4 N5 P6, black negative)."""
import numpy as np

K, Q, R, B, N, P = 1, 2, 3, 4, 5, 6

# gym_chess/envs/chess_v2.py:59-68 (DEFAULT_BOARD)
DEFAULT_BOARD = np.array([-R, -N, -B, -Q, -K, -B, -N, -R] + [-P] * 8 + [0] * 32 + [P] * 8 + [R, N, B, Q, K, B, N, R], np.int8)


def _board(**pieces):
    b = np.zeros(64, np.int8)
    for sq, p in pieces.items():
        b[int(sq[1:])] = p
    return b


def endgame_boards():
    """BASELINE.json configs[4]: repetition-heavy endgames (few irreversible moves -> long repetition windows)"""
    return np.array([
        _board(s60=K, s4=-K),                        # K v K
        _board(s60=K, s59=R, s4=-K),                 # KR v K
        _board(s60=K, s58=B, s57=N, s4=-K),          # KBN v K
        _board(s60=K, s59=Q, s4=-K, s3=-R),          # KQ v KR
        _board(s60=K, s62=N, s4=-K, s1=-N),          # KN v KN
        _board(s60=K, s52=P, s4=-K, s12=-P),         # KP v KP (pawns stall on the last rank: no promotion, Q1)
        _board(s60=K, s61=B, s4=-K, s2=-B),          # KB v KB
    ], np.int8)
