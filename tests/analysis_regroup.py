"""Analysis helper (not a test; run by hand: python tests/analysis_regroup.py): would re-grouping envs into warps by game phase or
material pay?  Replays phase-spread random self-play on the CPU oracle, records every position, and sums -- per warp of 32 envs
and per step -- the trips of the divergent piece loops of the step kernel (max over the lanes of each per-type count, weighted
with the measured instructions per trip) for random grouping, for a device-wide sort by ply / material / own pawns redone every
L steps, and for a sort inside each block of 128 envs.  Result (4096 envs, 360 steps): 872 divergent instructions per warp and
step with random grouping; device-wide sort by ply every 64 / 32 / 16 steps 803 / 783 / 763; re-sorted every step by own pawns
643; inside blocks of 128 envs 830 -- i.e. 3-4 % of the kernel's 2 390 instructions per warp and step before the cost of the
sort and of the gathers, because most of the lane imbalance is the spread of piece counts AT a given ply, not the mix of plies."""
import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
E, T = 4096, 720
t0 = time.time()
b, p, r = O.harvest(2, 0, E, T, 1, threads=16)
print("harvest", b.shape, time.time() - t0)
b = b.reshape(E, T, 64); p = p.reshape(E, T)
# piece ids: K1 Q2 R3 B4 N5 P6, black negative; player +1 white / -1 black
own = b * p[:, :, None]          # own pieces positive
cnt = lambda code, sign: ((own == sign * code).sum(axis=2)).astype(np.int16)
feat = {"pawn": cnt(6, 1), "knight": cnt(5, 1), "king": cnt(1, 1), "rook": cnt(3, 1), "bishop": cnt(4, 1), "queen": cnt(2, 1),
        "eRQ": cnt(3, -1) + cnt(2, -1), "eBQ": cnt(4, -1) + cnt(2, -1)}
cost = {"pawn": 37, "knight": 30, "king": 30, "rook": 59, "bishop": 59, "queen": 85, "eRQ": 46 * 0.6, "eBQ": 46 * 0.6}
total = (b != 0).sum(axis=2)
# ply index within the episode: detect resets (total pieces jumps to 32)
ply = np.zeros((E, T), np.int32)
for t in range(1, T):
    reset = (total[:, t] == 32) & (total[:, t - 1] < 32) | ((total[:, t] == 32) & (ply[:, t - 1] > 40))
    ply[:, t] = np.where(reset, 0, ply[:, t - 1] + 1)
rng = np.random.default_rng(0)
off = rng.integers(0, 301, E)          # phase spread
S = 360                                # global steps simulated
idx = lambda s: off + s                # env i shows its own step off_i + s
def run(L, keyfn):
    tot = 0.0; n = 0
    for w0 in range(0, S, L):
        ii = idx(w0)
        key = keyfn(ii)
        order = np.argsort(key, kind="stable")
        groups = order.reshape(-1, 32)
        for s in range(w0, min(S, w0 + L)):
            jj = idx(s)
            for k, c in cost.items():
                v = feat[k][np.arange(E), jj][groups]       # [warps, 32]
                tot += c * v.max(axis=1).sum()
            n += groups.shape[0]
    return tot / n
ar = np.arange(E)
rand = run(64, lambda ii: rng.random(E))
print("random grouping: divergent instr per warp-step %.0f" % rand)
for L in (1, 16, 32, 64):
    byply = run(L, lambda ii: ply[ar, ii])
    bymat = run(L, lambda ii: total[ar, ii])
    bypawn = run(L, lambda ii: feat["pawn"][ar, ii] * 100 + total[ar, ii])
    print("L=%2d  by ply %.0f  by material %.0f  by own pawns %.0f" % (L, byply, bymat, bypawn))
# block-level (128 envs) sort
def run_block(L, keyfn):
    tot = 0.0; n = 0
    for w0 in range(0, S, L):
        ii = idx(w0); key = keyfn(ii)
        blocks = np.arange(E).reshape(-1, 128)
        order = np.take_along_axis(blocks, np.argsort(key[blocks], axis=1, kind="stable"), axis=1)
        groups = order.reshape(-1, 32)
        for s in range(w0, min(S, w0 + L)):
            jj = idx(s)
            for k, c in cost.items():
                v = feat[k][np.arange(E), jj][groups]
                tot += c * v.max(axis=1).sum()
            n += groups.shape[0]
    return tot / n
for L in (32, 64):
    print("block-level L=%d by ply %.0f by material %.0f" % (L, run_block(L, lambda ii: ply[ar, ii]), run_block(L, lambda ii: total[ar, ii])))
