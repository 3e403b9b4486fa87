"""`ChessEngine` -- drop-in for the reference's PyO3 class (src/lib.rs:1412-1512) backed by the CUDA library.

Same 4 methods, same dict / str wire format, same results; each call runs a batch of one position through the
host-buffer entry points of the C ABI (gcb_host_*).  `BatchedChessEngine` exposes the same operations over numpy
arrays of positions.  No CPU fallback: without the library or a GPU these raise.
"""
import numpy as np

from . import _lib
from ._lib import check
from .codec import ACTION_TO_STR, STR_TO_ACTION

WHITE, BLACK = "WHITE", "BLACK"
_RIGHT_KEYS = ("white_king_castle_is_possible", "white_queen_castle_is_possible", "black_king_castle_is_possible",
               "black_queen_castle_is_possible")


def _player(p):
    # lib.rs:424-441: an unknown colour sets an exception but the call goes on as White
    return -1 if p == BLACK else 1


class BatchedChessEngine:
    """numpy wire format: boards int8[n,64|8,8], players int8[n] (+1/-1), rights uint8[n,4]"""

    @staticmethod
    def _prep(boards, players, rights):
        boards = np.ascontiguousarray(np.asarray(boards, np.int8).reshape(-1, 64))
        n = boards.shape[0]
        players = np.ascontiguousarray(np.broadcast_to(np.asarray(players, np.int8), (n,)))
        rights = np.ascontiguousarray(np.broadcast_to(np.asarray(rights, np.uint8), (n, 4)))
        return n, boards, players, rights

    def get_possible_moves(self, boards, players, rights, attack=False, castles_only=False, stride=256):
        """-> (actions uint16[n,stride] in reference order, counts int32[n], in_check uint8[n])"""
        n, boards, players, rights = self._prep(boards, players, rights)
        out, cnt, chk = np.zeros((n, stride), np.uint16), np.zeros(n, np.int32), np.zeros(n, np.uint8)
        check(_lib.lib().gcb_host_get_possible_moves(n, boards.ctypes.data, players.ctypes.data, rights.ctypes.data,
                                                     int(bool(attack)), int(bool(castles_only)), out.ctypes.data, stride,
                                                     cnt.ctypes.data, chk.ctypes.data))
        return out, cnt, chk

    def next_state(self, boards, players, rights, actions):
        """-> (boards int8[n,64], rights uint8[n,4], checks uint8[n,2], reward int32[n], status int8[n]: 0 ok, 1 ok with both
        kings in check afterwards (Q19), -1 empty from-square, -2 bad action code)"""
        n, boards, players, rights = self._prep(boards, players, rights)
        actions = np.ascontiguousarray(np.asarray(actions, np.int32).reshape(n))
        ob, orr, oc = np.zeros((n, 64), np.int8), np.zeros((n, 4), np.uint8), np.zeros((n, 2), np.uint8)
        rew, st = np.zeros(n, np.int32), np.zeros(n, np.int8)
        check(_lib.lib().gcb_host_next_state(n, boards.ctypes.data, players.ctypes.data, rights.ctypes.data,
                                             actions.ctypes.data, ob.ctypes.data, orr.ctypes.data, oc.ctypes.data,
                                             rew.ctypes.data, st.ctypes.data))
        return ob, orr, oc, rew, st

    def update_state(self, boards, rights):
        """-> (rights uint8[n,4] masked by king presence, checks uint8[n,2])"""
        n, boards, _, rights = self._prep(boards, 1, rights)
        orr, oc = np.zeros((n, 4), np.uint8), np.zeros((n, 2), np.uint8)
        check(_lib.lib().gcb_host_update_state(n, boards.ctypes.data, rights.ctypes.data, orr.ctypes.data, oc.ctypes.data))
        return orr, oc


class ChessEngine:
    """`from gym_chess import ChessEngine` replacement: dict in, dict out (lib.rs:1422-1511)."""

    def __init__(self):
        self._b = BatchedChessEngine()

    @staticmethod
    def _wire(state):
        board = np.asarray(state["board"], np.int8).reshape(1, 64)
        rights = np.array([[bool(state[k]) for k in _RIGHT_KEYS]], np.uint8)
        return board, rights

    @staticmethod
    def _dict(board, rights, checks, player):
        return dict(white_king_castle_is_possible=bool(rights[0]), white_queen_castle_is_possible=bool(rights[1]),
                    black_king_castle_is_possible=bool(rights[2]), black_queen_castle_is_possible=bool(rights[3]),
                    white_king_is_checked=bool(checks[0]), black_king_is_checked=bool(checks[1]),
                    board=[[int(v) for v in row] for row in board.reshape(8, 8)], current_player=player)

    def next_state(self, state, player, move):
        board, rights = self._wire(state)
        p = _player(player)
        ob, orr, oc, rew, st = self._b.next_state(board, p, rights, [STR_TO_ACTION[move]])
        if st[0] == -1:
            raise RuntimeError("Bad move - piece is empty !")  # the reference panics (lib.rs:693-695)
        if st[0] < 0:
            raise ValueError("bad move %r" % (move,))
        if st[0] == 1:
            # both kings in check after the move: the reference prints, sets an exception with PyErr::restore and still
            # returns Ok(...) (lib.rs:1442-1446, Q19) -- which CPython turns into this SystemError at the call site
            raise SystemError("<method 'next_state' of 'ChessEngine' objects> returned a result with an exception set "
                              "(Both Kings are in check: this position is impossible)")
        return self._dict(ob[0], orr[0], oc[0], BLACK if p > 0 else WHITE), int(rew[0])

    def get_possible_moves(self, state, player, attack=False):
        board, rights = self._wire(state)
        out, cnt, _ = self._b.get_possible_moves(board, _player(player), rights, attack=attack, stride=2048)
        return [ACTION_TO_STR[a] for a in out[0, : cnt[0]]]

    def get_castle_moves(self, state, player):
        board, rights = self._wire(state)
        out, cnt, _ = self._b.get_possible_moves(board, _player(player), rights, castles_only=True, stride=2048)
        return [ACTION_TO_STR[a] for a in out[0, : cnt[0]]]

    def update_state(self, state):
        board, rights = self._wire(state)
        orr, oc = self._b.update_state(board, rights)
        return self._dict(board[0], orr[0], oc[0], state["current_player"])
