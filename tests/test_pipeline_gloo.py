"""Layer-pipeline tick schedule on CPU: world_size 2 and 4 over gloo with a fake stage, checked against a
sequential single-process evaluation of the same per-sequence recurrences -- in both hand-off modes of
`RingPipeline.tick`: the grouped exchange per tick, and the mailbox order (wait -> step -> send, per-slot sequence
numbers) that the GPU runs use over NVLink peer memory (`parallel.PeerMailbox`)."""
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


class GlooMailbox:
    """CPU stand-in for parallel.PeerMailbox with the same interface and the same sequence-number rule: `send` posts the
    payload to the next rank (tag = slot), `wait_into` blocks on the upstream's k-th send for that slot -- except on
    stage 0, whose first wait per slot passes without a send and leaves the buffer alone (the prefill's token)."""

    def __init__(self, rank, world, S):
        self.rank, self.nxt, self.prv = rank, (rank + 1) % world, (rank - 1) % world
        self.waits = [-1 if rank == 0 else 0] * S
        self.pending = []

    def send(self, slot, src):
        self.pending.append(dist.isend(src.clone(), self.nxt, tag=slot))

    def wait_into(self, slot, dst):
        self.waits[slot] += 1
        if self.waits[slot] == 0:
            return
        dist.recv(dst, self.prv, tag=slot)

    def drain(self):
        for w in self.pending:
            w.wait()


def _worker(rank, world, port, n_ticks, q, mode="collective"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = world
    st = FakeStage(rank, world, S)
    first = [7 + 3 * s for s in range(S)]
    if st.first:
        for s in range(S):
            st.tok[s] = torch.tensor([first[s]])      # what the prefill phase leaves behind
    mailbox = GlooMailbox(rank, world, S) if mode == "mailbox" else None
    pipe = RingPipeline(st, rank, world, S, mailbox=mailbox)
    emitted = 0
    for _ in range(n_ticks):
        emitted += bool(pipe.tick())
    if st.last:
        q.put((st.emitted, emitted))
    if mailbox is not None:
        # the upstream's send of its last tick has no tick left to consume it: receive it, then retire the own sends
        prv = (rank - 1) % world
        s_last = (n_ticks - 1 - prv) % S
        dist.recv(st.in_buffer(s_last).clone(), prv, tag=s_last)
        mailbox.drain()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "collective"), (4, "collective"), (2, "mailbox"), (4, "mailbox")])
def test_ring_pipeline_matches_sequential(world, mode):
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_ticks = world - 1 + 3 * world                      # fill + 3 tokens per sequence
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_ticks, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    emitted, count = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ref = _sequential(world, world, [7 + 3 * s for s in range(world)], 3)
    assert emitted == ref
    assert count == 3 * world                            # one token per tick once the pipe is full
