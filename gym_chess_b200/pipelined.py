"""PipelinedChessEnv -- the envs of one device as K shards stepped alternately through page-locked host buffers.

The host-facing fast path (what a vector-env wrapper of the reference's `ChessEnvV2` would be built on): actions and
results live in page-locked host memory that the step kernel reads / writes in place through PCIe, as 16-bit records
(uint16 action or random word in; uint16 `reward int8 | flags << 8 | done << 15` out, include/gymchess_b200.h), and
the shards are stepped asynchronously: while the device steps one shard the host has the other's results and writes its
next actions (EnvPool-style send / recv), so neither the launch / completion latency of a synchronous `step()` nor the
last-wave tail of a kernel is paid.  Global env ids (Philox counters) are those of one env set of `num_envs`.

    pipe = PipelinedChessEnv(524288, shards=2, opponent="none")
    pipe.send_words(0); pipe.send_words(1)          # or send_actions(k) after writing pipe.inputs[k]
    while training:
        for k in range(pipe.shards):
            rec = pipe.recv(k)                      # uint16 view of shard k's results (valid until its next send)
            ...                                      # consume, write pipe.inputs[k][:] = next actions
            pipe.send_actions(k)
"""
import numpy as np
import torch

from .batched_env import BatchedChessEnv


class PipelinedChessEnv:
    def __init__(self, num_envs, shards=2, device=0, env_id_offset=0, observe=False, mask=False, **env_kwargs):
        """observe=True: every step also brings the observation back -- the four 64-bit bit-planes of the board (32 B per env,
        the resident form of `state["board"]`; piece code = reference id in planes t0..t2, colour plane `white`) are copied to
        page-locked host memory behind the step, on the shard's stream: `observations[k]` = (int64 [H, 2] t0|t1, int64 [H, 2]
        t2|white).  mask=True: likewise the bit mask of possible_actions (int64 [H, 66], written by the step kernel itself,
        gcb_env_step_mask_output): `masks[k]`.  Both are valid after recv(k), like the result records."""
        if shards < 1 or num_envs % shards:
            raise ValueError("num_envs must be a multiple of shards")
        self.num_envs, self.shards, self.shard_envs = int(num_envs), int(shards), int(num_envs) // int(shards)
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        H = self.shard_envs
        self.envs = [BatchedChessEnv(H, device=device, env_id_offset=env_id_offset + k * H, **env_kwargs) for k in range(shards)]
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(shards)]
        self._in = [torch.zeros(H, dtype=torch.int16).pin_memory() for _ in range(shards)]
        self._out = [torch.zeros(H, dtype=torch.int16).pin_memory() for _ in range(shards)]
        # the caller writes actions / random words here (uint16 views of the page-locked buffers) ...
        self.inputs = [t.numpy().view(np.uint16) for t in self._in]
        # ... and reads the result records here after recv()
        self.results = [t.numpy().view(np.uint16) for t in self._out]
        self._in_flight = [False] * shards
        self.observe, self.mask = bool(observe), bool(mask)
        self.observations, self.masks, self._planes, self._dmask = [], [], [], []
        for e in self.envs:
            if self.observe:
                self._planes.append(e.planes())
                self.observations.append(tuple(torch.zeros((H, 2), dtype=torch.int64).pin_memory() for _ in range(2)))
            if self.mask:
                with torch.cuda.device(self.device):
                    m = torch.zeros((H, 66), dtype=torch.int64, device=self.device)
                e.set_mask_output(m)
                self._dmask.append(m)
                self.masks.append(torch.zeros((H, 66), dtype=torch.int64).pin_memory())

    def _after_step(self, k):
        # device -> host copies of the step's other outputs, enqueued behind the step kernel on the shard's stream
        if self.observe or self.mask:
            with torch.cuda.stream(self.streams[k]):
                if self.observe:
                    for dst, src in zip(self.observations[k], self._planes[k]):
                        dst.copy_(src, non_blocking=True)
                if self.mask:
                    self.masks[k].copy_(self._dmask[k], non_blocking=True)
        self._in_flight[k] = True

    def send_actions(self, k, src=None):
        """enqueue ChessEnvV2.step(action) of every env of shard k with the actions in inputs[k] (action codes < 4101), or in
        `src`: any other page-locked int16/uint16 buffer of shard_envs entries (e.g. where a policy's output already lies),
        read in place -- it must stay untouched until recv(k)"""
        self.envs[k].step_packed(self._in[k] if src is None else src, self._out[k], stream=self.streams[k])
        self._after_step(k)

    def send_words(self, k, src=None):
        """enqueue a step in which env i plays possible_actions[(word[i] * n_legal) >> 16] (a uniform random legal action for
        uniform 16-bit words; RESIGN when it has no legal move); words from inputs[k] or from a page-locked `src`"""
        self.envs[k].step_index_packed(self._in[k] if src is None else src, self._out[k], stream=self.streams[k])
        self._after_step(k)

    def recv(self, k):
        """block until shard k's step has finished; returns results[k] (uint16 records, see unpack)"""
        if self._in_flight[k]:
            self.envs[k].wait(stream=self.streams[k])
            self._in_flight[k] = False
        return self.results[k]

    unpack = staticmethod(BatchedChessEnv.unpack_result)

    def burn_in(self, steps, dephase=False):
        """`steps` sampled self-play steps of every shard (on-device draws), e.g. to mix game phases before measuring;
        dephase=True first spreads the episode phases (BatchedChessEnv.dephase)"""
        for k, e in enumerate(self.envs):
            with torch.cuda.stream(self.streams[k]):
                if dephase:
                    e.dephase()
                e.step_sampled(steps)
        for k in range(self.shards):
            self.streams[k].synchronize()

    def stats(self):
        tot = None
        for e in self.envs:
            s = e.stats()
            tot = s if tot is None else {k: tot[k] + s[k] for k in s}
        return tot

    def close(self):
        for e in self.envs:
            e.close()
