#!/usr/bin/env python
"""The FIXED 1,048,576-position set of BASELINE.json configs[1] / SURVEY.md 8(d) "Config 2", committed by recipe.

    python tests/golden/make_positions_1m.py        # regenerates the set with the oracle and rewrites positions_1m.sha256

The set is a pure function of this file (seeded numpy generators, seeded self-play), so it is not stored: the tests and
`bench.py`'s movegen leg rebuild it and check its SHA-256 against the committed `positions_1m.sha256`.

  * ~90 %: positions of random self-play games from the start position (seed 1, env ids 0..9599, 300 steps each with
    auto-reset), recorded so that every ply index is equally likely -- all game phases, ~6 % in check.  `selfplay` is a
    callable: the tests pass the oracle's harvest (oracle.harvest), bench.py passes the CUDA env's (harvest_gpu below);
    the draws are Philox counters of (seed, env id, episode, step), so both give the SAME positions -- which
    tests/test_gpu_parity.py::test_config2_fixed_positions also asserts.
  * ~10 %: crafted positions, both sides to move:
      - every board the reference's own v2 tests use (gym_chess/test/v2/*.py: test_basic_moves, test_capture_moves,
        test_king_moves, test_squares_under_attack, test_castle_moves, test_run_moves; taken from the engine calls
        logged in reference_v2_tests.json) under all 16 castle-right combinations;
      - the 1,935 harvested / edge positions of positions.json.gz;
      - castle shapes: rooks and king on their home squares with random company, attackers aimed at the transit
        squares (castle through / out of / into an attacked square), every rights combination (the OR rule, Q4), white
        pieces on rank 8 (the white-id test of the black branch, Q3);
      - kings on the ray of an enemy slider (the x-ray hole, Q6), kingless boards (Q7), several kings of one colour
        (Q15), pawns on rows 0 and 7 (dead promotion, Q1), pawns on their start row behind a blocker (the jump, Q13),
        pawn attacks onto the own king (Q14), single and double checks with pinned pieces around the king;
      - random piece soups.
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TOTAL = 1 << 20
SELFPLAY_SEED, SELFPLAY_ENVS, SELFPLAY_STEPS, SELFPLAY_EVERY = 1, 9600, 300, 3   # -> 960,000 positions
K, Q, R, B, N, P = 1, 2, 3, 4, 5, 6


def _sq(name):
    return (8 - int(name[1])) * 8 + "abcdefgh".index(name[0])


def _reference_test_boards():
    with open(os.path.join(HERE, "reference_v2_tests.json")) as f:
        rep = json.load(f)
    seen, out = set(), []
    for rec in rep:
        for c in rec.get("calls", []):
            for key in ("state",):
                b = tuple(c[key]["board"])
                if b not in seen:
                    seen.add(b), out.append(b)
            if "board_after" in c and tuple(c["board_after"]) not in seen:
                seen.add(tuple(c["board_after"])), out.append(tuple(c["board_after"]))
    return np.array(out, np.int8)


def _golden_positions():
    with gzip.open(os.path.join(HERE, "positions.json.gz")) as f:
        pos = json.loads(f.read())
    return np.array([p["board"] for p in pos], np.int8), np.array([p["rights"] for p in pos], np.uint8)


def _sprinkle(rng, b, k, ids, avoid=()):
    """k random pieces from `ids` on empty squares of b (not on `avoid`)"""
    free = [s for s in np.nonzero(b == 0)[0] if s not in avoid]
    for s in rng.choice(free, size=min(k, len(free)), replace=False):
        b[s] = rng.choice(ids)


def _crafted(rng, n):
    """n crafted boards (int8[n,64]) cycling through the edge-case families"""
    out = np.zeros((n, 64), np.int8)
    men = [-P, -N, -B, -R, -Q, P, N, B, R, Q]
    for i in range(n):
        b = out[i]
        fam = i % 10
        if fam == 0:    # white castle shapes, attackers on the transit squares
            b[_sq("e1")], b[_sq("a1")], b[_sq("h1")] = K, rng.choice([R, R, R, 0, -R]), rng.choice([R, R, R, 0, N])
            b[rng.randint(0, 16)] = -K
            _sprinkle(rng, b, rng.randint(0, 10), men, avoid=range(56, 64))
            if rng.rand() < 0.4:
                b[rng.choice([57, 58, 59, 61, 62])] = rng.choice([N, B, -N, Q])
            if rng.rand() < 0.7:  # a slider / knight aimed at c1..g1
                f = rng.randint(2, 7)
                b[rng.randint(1, 6) * 8 + f] = rng.choice([-R, -Q])
        elif fam == 1:  # black castle shapes, incl. WHITE rooks / king on rank 8 (Q3)
            b[_sq("e8")] = rng.choice([-K, K])
            b[_sq("a8")], b[_sq("h8")] = rng.choice([-R, R, 0]), rng.choice([-R, R, 0])
            b[rng.randint(48, 64)] = K if b[_sq("e8")] == -K else -K
            _sprinkle(rng, b, rng.randint(0, 10), men, avoid=range(0, 8))
        elif fam == 2:  # king on the ray of an enemy slider (Q6), blockers / pinned pieces in between
            ksq = rng.randint(64)
            kr, kc = divmod(ksq, 8)
            side = rng.choice([1, -1])
            b[ksq] = K * side
            dr, dc = [(0, 1), (1, 0), (1, 1), (1, -1), (0, -1), (-1, 0), (-1, -1), (-1, 1)][rng.randint(8)]
            ray = []
            r_, c_ = kr + dr, kc + dc
            while 0 <= r_ < 8 and 0 <= c_ < 8:
                ray.append(r_ * 8 + c_)
                r_, c_ = r_ + dr, c_ + dc
            if ray:
                far = ray[rng.randint(len(ray))]
                b[far] = -side * (rng.choice([R, Q]) if dr == 0 or dc == 0 else rng.choice([B, Q]))
                between = ray[: ray.index(far)]
                for s in between:
                    if rng.rand() < 0.25:
                        b[s] = rng.choice([side, -side]) * rng.choice([P, N, B, R, Q])
            free = np.nonzero(b == 0)[0]
            b[rng.choice(free)] = -K * side
            _sprinkle(rng, b, rng.randint(0, 12), men)
        elif fam == 3:  # kingless (Q7): one or both kings missing
            _sprinkle(rng, b, rng.randint(2, 20), men)
            if rng.rand() < 0.6:
                b[rng.choice(np.nonzero(b == 0)[0])] = rng.choice([K, -K])
        elif fam == 4:  # several kings of one colour (Q15)
            _sprinkle(rng, b, rng.randint(2, 14), men)
            for _ in range(rng.randint(2, 5)):
                b[rng.choice(np.nonzero(b == 0)[0])] = rng.choice([K, K, -K])
        elif fam == 5:  # pawns on rows 0 / 7 (Q1) and on their start rows behind a blocker (Q13)
            b[rng.randint(16, 48)] = K
            b[rng.choice(np.nonzero(b == 0)[0])] = -K
            for _ in range(rng.randint(1, 6)):
                b[rng.choice([rng.randint(0, 8), rng.randint(56, 64)])] = rng.choice([P, -P])
            for _ in range(rng.randint(1, 5)):
                c = rng.randint(8)
                if rng.rand() < 0.5:
                    b[48 + c], b[40 + c] = P, rng.choice(men)
                    b[32 + c] = rng.choice([0, 0, -P, N])
                else:
                    b[8 + c], b[16 + c] = -P, rng.choice(men)
                    b[24 + c] = rng.choice([0, 0, P, -N])
            _sprinkle(rng, b, rng.randint(0, 8), men)
        elif fam == 6:  # double check: a knight and a slider (or two sliders) on the king, defenders around
            ksq = rng.randint(8, 56)
            kr, kc = divmod(ksq, 8)
            side = rng.choice([1, -1])
            b[ksq] = K * side
            jumps = [(kr + a, kc + c) for a, c in ((-2, -1), (-2, 1), (2, -1), (2, 1), (-1, -2), (-1, 2), (1, -2), (1, 2))
                     if 0 <= kr + a < 8 and 0 <= kc + c < 8]
            jr, jc = jumps[rng.randint(len(jumps))]
            b[jr * 8 + jc] = -side * N
            line = [s for s in list(range(kr * 8, kr * 8 + 8)) + list(range(kc, 64, 8)) if s != ksq and b[s] == 0]
            b[rng.choice(line)] = -side * rng.choice([R, Q])
            b[rng.choice(np.nonzero(b == 0)[0])] = -K * side
            _sprinkle(rng, b, rng.randint(0, 10), [side * x for x in (P, N, B, R, Q)] + [-side * P])
        elif fam == 7:  # pawn attacks next to kings (Q14 / Q22): pawns and both kings packed together
            c0 = rng.randint(1, 7)
            r0 = rng.randint(1, 7)
            cells = [(r0 + a) * 8 + c0 + c for a in (-1, 0, 1) for c in (-1, 0, 1)]
            pick = rng.permutation(cells)
            b[pick[0]], b[pick[1]] = K, -K
            for s in pick[2: 2 + rng.randint(1, 6)]:
                b[s] = rng.choice([P, -P, P, -P, N, -B])
            _sprinkle(rng, b, rng.randint(0, 6), men)
        elif fam == 8:  # slider-heavy soups
            _sprinkle(rng, b, rng.randint(3, 22), [-B, -R, -Q, B, R, Q, K, -K])
        else:           # anything goes
            _sprinkle(rng, b, rng.randint(1, 28), men + [K, -K])
    return out


def build(selfplay, verbose=False):
    """-> (boards int8[N,64], players int8[N], rights uint8[N,4]), N = 1,048,576.
    selfplay(seed, n_envs, nsteps, every) -> (boards, players, rights) of the self-play share."""
    sb, sp, sr = selfplay(SELFPLAY_SEED, SELFPLAY_ENVS, SELFPLAY_STEPS, SELFPLAY_EVERY)
    assert len(sb) == SELFPLAY_ENVS * SELFPLAY_STEPS // SELFPLAY_EVERY, len(sb)
    rng = np.random.RandomState(20261018)
    parts_b, parts_r = [], []
    ref = _reference_test_boards()                       # the reference's own test boards x 16 rights combinations
    combos = np.array([[(m >> k) & 1 for k in range(4)] for m in range(16)], np.uint8)
    parts_b.append(np.repeat(ref, 16, axis=0)), parts_r.append(np.tile(combos, (len(ref), 1)))
    gb, gr = _golden_positions()
    parts_b.append(gb), parts_r.append(gr)
    n_fixed = sum(len(x) for x in parts_b)
    n_crafted_half = (N_TOTAL - len(sb)) // 2            # every crafted board appears with White and with Black to move
    cb = _crafted(rng, n_crafted_half - n_fixed)
    parts_b.append(cb), parts_r.append(rng.randint(0, 2, size=(len(cb), 4)).astype(np.uint8))
    hb, hr = np.concatenate(parts_b), np.concatenate(parts_r)
    boards = np.concatenate([sb, hb, hb])
    players = np.concatenate([sp, np.ones(len(hb), np.int8), -np.ones(len(hb), np.int8)])
    rights = np.concatenate([sr, hr, hr])
    assert len(boards) == N_TOTAL, len(boards)
    perm = np.random.RandomState(7).permutation(N_TOTAL)  # crafted and self-play positions share warps
    boards, players, rights = np.ascontiguousarray(boards[perm]), np.ascontiguousarray(players[perm]), np.ascontiguousarray(rights[perm])
    if verbose:
        print("self-play %d, reference-test boards %d x 16 x 2, golden %d x 2, generated %d x 2" % (len(sb), len(ref), len(gb), len(cb)))
    return boards, players, rights


def digest(boards, players, rights):
    h = hashlib.sha256()
    for a in (boards, players, rights):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def committed_digest():
    with open(os.path.join(HERE, "positions_1m.sha256")) as f:
        return f.read().split()[0]


def harvest_oracle(seed, n_envs, nsteps, every):
    from oracle import oracle as orc

    return orc.harvest(seed, 0, n_envs, nsteps, every, threads=os.cpu_count() or 1)


def harvest_gpu(seed, n_envs, nsteps, every, device=0):
    """the same positions from the CUDA env (bench.py: no oracle on the measured side): env ids, seeds and Philox counters
    are those of the oracle's harvest, so the games -- and the recorded positions -- are identical"""
    import torch

    from gym_chess_b200 import BatchedChessEnv

    env = BatchedChessEnv(n_envs, opponent="none", seed=seed, device=device, auto_reset=True)
    ids = torch.arange(n_envs, device=env.device)
    bs, ps, rs, key = [], [], [], []
    for t in range(nsteps):
        sel = (ids % every) == (t % every)
        info = env.info_tensor()[sel]
        bs.append(env.observe().reshape(n_envs, 64)[sel]), ps.append(info[:, 0].to(torch.int8)), rs.append(info[:, 1:5].to(torch.uint8))
        key.append(ids[sel] * nsteps + t)
        env.step_sampled(1)
    order = torch.argsort(torch.cat(key))                # the oracle records env by env
    out = tuple(torch.cat(x)[order].cpu().numpy() for x in (bs, ps, rs))
    env.close()
    return out


if __name__ == "__main__":
    b, p, r = build(harvest_oracle, verbose=True)
    d = digest(b, p, r)
    with open(os.path.join(HERE, "positions_1m.sha256"), "w") as f:
        f.write("%s  positions_1m (boards int8[1048576,64] | players int8 | rights uint8[.,4], tests/golden/make_positions_1m.py)\n" % d)
    from oracle import oracle as orc

    out, cnt = orc.movegen_batch(b, p, r, False, stride=256, threads=os.cpu_count() or 1)
    orr, oc = orc.update_state_batch(b, r)
    chk = np.where(p > 0, oc[:, 0], oc[:, 1])
    print("sha256", d)
    print("mean legal %.2f, max %d, in check %.3f, no legal move %.4f, white to move %.3f" % (
        cnt.mean(), cnt.max(), chk.mean(), (cnt == 0).mean(), (p > 0).mean()))
