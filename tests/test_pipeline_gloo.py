"""Layer-pipeline tick schedule on CPU: world_size 2 and 4 over gloo with a fake stage, checked against a
sequential single-process evaluation of the same per-sequence recurrences."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quip_for_all_b200.parallel import RingPipeline, partition_layers


def test_partition_layers():
    assert [len(r) for r in partition_layers(32, 8)] == [4] * 8
    assert [len(r) for r in partition_layers(80, 8)] == [10] * 8
    parts = partition_layers(32, 3)
    assert [len(r) for r in parts] == [11, 11, 10] and parts[0][0] == 0 and parts[-1][-1] == 31
    assert [list(r) for r in partition_layers(2, 2)] == [[0], [1]]


class FakeStage:
    """stage r:  hidden_out = hidden_in * (r + 2) + slot ; stage 0 embeds tok -> tok + 0.5 ; the last
    stage 'samples' tok = floor(hidden) mod 1000.  Everything in float64 / int64 on CPU."""

    def __init__(self, rank, world, S):
        self.r, self.world, self.S = rank, world, S
        self.first, self.last = rank == 0, rank == world - 1
        self.tok = [torch.zeros(1, dtype=torch.long) for _ in range(S)]
        self.hin = [torch.zeros(4, dtype=torch.float64) for _ in range(S)]
        self.hout = [torch.zeros(4, dtype=torch.float64) for _ in range(S)]
        self.emitted = [[] for _ in range(S)]
        self._scratch = torch.zeros(1, dtype=torch.long) if self.first else torch.zeros(4, dtype=torch.float64)

    @staticmethod
    def stage_fn(r, s, h):
        return h * (r + 2) + s

    def step(self, s):
        h = (self.tok[s].double() + 0.5).expand(4).clone() if self.first else self.hin[s]
        h = self.stage_fn(self.r, s, h)
        if self.last:
            self.tok[s] = (h[:1].floor().long() % 1000)
            self.emitted[s].append(int(self.tok[s]))
        else:
            self.hout[s] = h

    def out_buffer(self, s):
        return self.tok[s] if self.last else self.hout[s]

    def in_buffer(self, s):
        return self.tok[s] if self.first else self.hin[s]

    def scratch_in(self):
        return self._scratch


def _sequential(world, S, first_tokens, n_tokens):
    out = []
    for s in range(S):
        tok, seq = first_tokens[s], []
        for _ in range(n_tokens):
            h = torch.full((4,), tok + 0.5, dtype=torch.float64)
            for r in range(world):
                h = FakeStage.stage_fn(r, s, h)
            tok = int(h[0].floor().long() % 1000)
            seq.append(tok)
        out.append(seq)
    return out


def _worker(rank, world, port, n_ticks, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = world
    st = FakeStage(rank, world, S)
    first = [7 + 3 * s for s in range(S)]
    if st.first:
        for s in range(S):
            st.tok[s] = torch.tensor([first[s]])      # what the prefill phase leaves behind
    pipe = RingPipeline(st, rank, world, S)
    emitted = 0
    for _ in range(n_ticks):
        emitted += bool(pipe.tick())
    if st.last:
        q.put((st.emitted, emitted))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_ring_pipeline_matches_sequential(world):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_ticks = world - 1 + 3 * world                      # fill + 3 tokens per sequence
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_ticks, q)) for r in range(world)]
    for p in procs:
        p.start()
    emitted, count = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _sequential(world, world, [7 + 3 * s for s in range(world)], 3)
    assert emitted == ref
    assert count == 3 * world                            # one token per tick once the pipe is full
