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
    def __init__(self, num_envs, shards=2, device=0, env_id_offset=0, **env_kwargs):
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

    def send_actions(self, k, src=None):
        """enqueue ChessEnvV2.step(action) of every env of shard k with the actions in inputs[k] (action codes < 4101), or in
        `src`: any other page-locked int16/uint16 buffer of shard_envs entries (e.g. where a policy's output already lies),
        read in place -- it must stay untouched until recv(k)"""
        self.envs[k].step_packed(self._in[k] if src is None else src, self._out[k], stream=self.streams[k])
        self._in_flight[k] = True

    def send_words(self, k, src=None):
        """enqueue a step in which env i plays possible_actions[(word[i] * n_legal) >> 16] (a uniform random legal action for
        uniform 16-bit words; RESIGN when it has no legal move); words from inputs[k] or from a page-locked `src`"""
        self.envs[k].step_index_packed(self._in[k] if src is None else src, self._out[k], stream=self.streams[k])
        self._in_flight[k] = True

    def recv(self, k):
        """block until shard k's step has finished; returns results[k] (uint16 records, see unpack)"""
        if self._in_flight[k]:
            self.envs[k].wait(stream=self.streams[k])
            self._in_flight[k] = False
        return self.results[k]

    unpack = staticmethod(BatchedChessEnv.unpack_result)

    def burn_in(self, steps):
        """`steps` sampled self-play steps of every shard (on-device draws), e.g. to mix game phases before measuring"""
        for k, e in enumerate(self.envs):
            with torch.cuda.stream(self.streams[k]):
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
