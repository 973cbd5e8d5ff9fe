"""Multi-GPU layer pipeline: one process per GPU, whole decoder layers per stage, activations shipped
between stages with NCCL point-to-point over NVLink (torch.distributed).

Why layers and not intra-linear shards: the Hadamard rotations of a QuantLinear span its full input
and output dimensions, so a linear cannot be row/column sharded (the reference says the same,
README.md:84); the reference's only multi-device facility keeps whole decoder blocks on one device
(`no_split_module_classes`, quantizer.py:180-191, :831).  There is no collective on the data path --
only the [1, 1, hidden] fp16 hand-off (8-16 KiB) per stage boundary and the 8-byte token id back to
stage 0.

Schedule.  P stages keep S = P independent bs=1 sequences in flight.  Time is cut into ticks; at tick
t stage r runs one decode step of sequence (t - r) mod S, then ALL ranks do one ring exchange
(rank r -> r+1, last -> 0) issued as a single grouped isend/irecv, which is deadlock-free under strict
rendezvous semantics.  The token sampled by the last stage at tick t reaches stage 0 for tick t+1, which
is exactly when stage 0 is due to run that sequence again.  One token leaves the pipeline per tick.
"""
import os
import time
from typing import List

import torch
import torch.distributed as dist


def partition_layers(n_layers: int, world: int) -> List[range]:
    """Contiguous, near-equal slices of the decoder layers (earlier stages get the remainder)."""
    base, rem = divmod(n_layers, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < rem else 0)
        out.append(range(start, start + n))
        start += n
    return out


class PeerMailbox:
    """Device-side hand-off over NVLink peer memory (csrc/handoff.cu).  Every rank owns a mailbox with one slot per
    in-flight sequence ([payload | flag]); the upstream rank maps it through CUDA IPC and stores into it.  No host
    synchronisation and no NCCL call on the tick path."""
    FLAG_OFF = 32768          # payload capacity per slot (>= 2 * hidden bytes of any supported model)
    SLOT = FLAG_OFF + 256

    def __init__(self, rank, world, n_slots, device, group=None):
        import ctypes
        from ._native import check, lib
        self.L, self.check, self.ct = lib(), check, ctypes
        self.rank, self.world, self.S, self.dev = rank, world, n_slots, device
        # Every rank takes part in every collective below whatever happens locally, and all ranks reach the same verdict:
        # a rank that failed to create / map a mailbox must not leave the others waiting in a barrier.
        own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        self.own = self.peer = None
        why = ""
        with torch.cuda.device(device):
            rc = self.L.quipb200_mailbox_create(self.SLOT * n_slots, ctypes.byref(own), handle)
        if rc == 0:
            self.own = own.value
        else:
            why = f"mailbox_create rc={rc}"
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw) if rc == 0 else b"", group=group)
        if rc == 0 and all(handles):
            peer = ctypes.c_void_p()
            with torch.cuda.device(device):
                rc = self.L.quipb200_mailbox_open(handles[(rank + 1) % world], ctypes.byref(peer))
            if rc == 0:
                self.peer = peer.value
            else:
                why = f"mailbox_open rc={rc}"
        elif rc == 0:
            rc, why = -1, "a peer could not create its mailbox"
        okt = torch.tensor([1 if rc == 0 else 0], device=device)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN, group=group)
        if okt.item() == 0:
            if self.peer:
                self.L.quipb200_mailbox_close(self.peer)
            dist.barrier(group=group)            # nobody frees a mailbox a neighbour still maps
            if self.own:
                self.L.quipb200_mailbox_destroy(self.own)
            self.own = self.peer = None
            raise RuntimeError("peer-memory mailboxes unavailable on at least one rank" + (f" ({why})" if why else ""))
        # per-slot sequence counters (device uint64): sends start at 0; waits start at 0, except on stage 0 whose first
        # run of a slot consumes the token left by the prefill (counter -1: the first wait passes and copies nothing)
        self.send_ctr = torch.zeros(n_slots, dtype=torch.int64, device=device)
        self.wait_ctr = torch.full((n_slots,), -1 if rank == 0 else 0, dtype=torch.int64, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(group=group)

    def _st(self):
        return self.ct.c_void_p(torch.cuda.current_stream().cuda_stream)

    def send(self, slot, src: torch.Tensor):
        nbytes = src.numel() * src.element_size()
        assert nbytes <= self.FLAG_OFF and nbytes % 8 == 0
        base = self.peer + slot * self.SLOT
        self.check(self.L.quipb200_handoff_send(src.data_ptr(), base, nbytes, base + self.FLAG_OFF,
                                                self.send_ctr.data_ptr() + 8 * slot, self._st()), "handoff_send")

    def wait_into(self, slot, dst: torch.Tensor):
        nbytes = dst.numel() * dst.element_size()
        base = self.own + slot * self.SLOT
        self.check(self.L.quipb200_handoff_wait(base + self.FLAG_OFF, self.wait_ctr.data_ptr() + 8 * slot, base,
                                                dst.data_ptr(), nbytes, self.err.data_ptr(), self._st()), "handoff_wait")

    def errors(self):
        return int(self.err.item())

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        if self.peer:
            self.L.quipb200_mailbox_close(self.peer)
            self.peer = None
        dist.barrier()
        if self.own:
            self.L.quipb200_mailbox_destroy(self.own)
            self.own = None


class RingPipeline:
    """Tick driver.  `stage` must provide, for slot s in [0, S):
         step(s)                      run one decode step of slot s in place
         out_buffer(s) / in_buffer(s) tensors sent to the next stage / received from the previous one
         scratch_in()                 a tensor shaped like in_buffer for discarded fill-phase traffic
    """

    def __init__(self, stage, rank: int, world: int, n_slots: int = None, group=None, mailbox: "PeerMailbox" = None):
        self.stage, self.rank, self.world = stage, rank, world
        self.mailbox = mailbox       # device-side hand-off (GPU runs); None: one grouped isend/irecv per tick
        self.S = n_slots or world
        assert world % self.S == 0 or self.S == world, "token feedback needs S | P"
        self.group = group
        self.nxt, self.prv = (rank + 1) % world, (rank - 1) % world
        self.t = 0

    def exchange(self, send_t: torch.Tensor, recv_t: torch.Tensor):
        if self.world == 1:
            recv_t.copy_(send_t)
            return
        ops = [dist.P2POp(dist.isend, send_t, self.nxt, self.group),
               dist.P2POp(dist.irecv, recv_t, self.prv, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def tick(self, pre_step=None):
        """`pre_step(slot)`: called after this tick's input has arrived and before the step consumes it (the e2e
        harness overrides the input token there)."""
        t, r, S = self.t, self.rank, self.S
        active = t >= r
        s = (t - r) % S
        if self.mailbox is not None:
            if active:
                self.mailbox.wait_into(s, self.stage.in_buffer(s))
                if pre_step is not None:
                    pre_step(s)
                self.stage.step(s)
                self.mailbox.send(s, self.stage.out_buffer(s))
            self.t += 1
            return active and r == self.world - 1
        if active:
            if pre_step is not None:
                pre_step(s)
            self.stage.step(s)
        s_next = (t + 1 - r) % S
        sender_active = t >= ((r - 1) % self.world)      # was my upstream neighbour active this tick?
        recv_t = self.stage.in_buffer(s_next) if sender_active else self.stage.scratch_in()
        self.exchange(self.stage.out_buffer(s), recv_t)
        self.t += 1
        return active and r == self.world - 1            # True when this tick emitted a token


# --------------------------------------------------------------------------------------------------
# the real stage: a slice of a random-init quantised Llama, one decode engine per in-flight sequence
# --------------------------------------------------------------------------------------------------
class LlamaStage:
    def __init__(self, model_name, codebook, rank, world, device, n_slots, cache_len, seed=0, use_graph=True,
                 **cfg_overrides):
        from .modeling import LlamaDecodeEngine, llama_config, make_random_quantized_llama
        cfg = llama_config(model_name, **cfg_overrides)
        self.layers = partition_layers(cfg.num_hidden_layers, world)[rank]
        self.first, self.last = rank == 0, rank == world - 1
        self.model = make_random_quantized_llama(cfg, codebook, seed=seed + rank, device=device,
                                                 layer_range=self.layers)
        self.engines = [LlamaDecodeEngine(self.model, max_cache_len=cache_len, first_stage=self.first,
                                          last_stage=self.last, use_cuda_graph=use_graph)
                        for _ in range(n_slots)]
        e = self.engines[0]
        self._scratch = torch.zeros_like(e.tok if self.first else e.hidden_in)
        self.hidden = cfg.hidden_size
        self.device = device

    def step(self, s):
        self.engines[s].step()

    def out_buffer(self, s):
        e = self.engines[s]
        return e.tok if self.last else e.hidden_out

    def in_buffer(self, s):
        e = self.engines[s]
        return e.tok if self.first else e.hidden_in

    def scratch_in(self):
        return self._scratch

    @torch.no_grad()
    def prefill_all(self, prompts, pipe: RingPipeline):
        """Sequential prefill of every slot through the stages (one-time, not timed)."""
        T = prompts[0].shape[1]
        world, r = pipe.world, pipe.rank
        hid = torch.zeros(1, T, self.hidden, dtype=torch.float16, device=self.device)
        tokbuf = torch.zeros(1, 1, dtype=torch.long, device=self.device)
        for s, ids in enumerate(prompts):
            e = self.engines[s]
            for stage in range(world):
                if stage == r:
                    out = e.prefill(ids.to(self.device) if self.first else hid)
                    if not self.last:
                        hid_out = out.contiguous()
                if stage < world - 1:                     # ship [1,T,H] from `stage` to `stage+1`
                    if r == stage:
                        dist.send(hid_out, stage + 1)
                    elif r == stage + 1:
                        dist.recv(hid, stage)
            # first token: last stage -> stage 0
            if world > 1:
                if self.last:
                    dist.send(e.tok, 0)
                elif self.first:
                    dist.recv(tokbuf, world - 1)
                    e.tok.copy_(tokbuf)
        for e in self.engines:
            e.capture()


def _pipeline_measure(a, model_name, world, rank, local, dev, clock_sampler_cls=None):
    """One layer-pipeline measurement of `model_name` on the already initialised process group.  Returns (on every
    rank) a dict with the device-timed and the end-to-end aggregate tokens/s."""
    S = world
    cache_len = a.cache_len or (a.prompt_len + (a.steps * 2 + a.warmup) // S * 2 + 4 * S + 32)
    stage = LlamaStage(model_name, a.codebook, rank, world, dev, S, cache_len, use_graph=not a.no_graph)
    mailbox = None
    if getattr(a, "handoff", "peer") == "peer":
        try:       # raises on EVERY rank or on none (the constructor agrees on the outcome through an all-reduce)
            mailbox = PeerMailbox(rank, world, S, dev)
        except RuntimeError as e:
            mailbox = None
            if rank == 0:
                print(f"[parallel] {e}; using NCCL p2p", flush=True)
    pipe = RingPipeline(stage, rank, world, S, mailbox=mailbox)
    g = torch.Generator().manual_seed(0)
    vocab = stage.model.config.vocab_size
    prompts = [torch.randint(0, vocab, (1, a.prompt_len), generator=g) for _ in range(S)]
    stage.prefill_all(prompts, pipe)
    # fill + warm-up ticks
    for _ in range(world - 1 + max(a.warmup, 3)):
        pipe.tick()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    clocks = clock_sampler_cls(local) if (clock_sampler_cls is not None and rank == 0) else None
    if clocks is not None:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        pipe.tick()                                       # one token leaves the pipeline per tick
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ck = clocks.stop() if clocks is not None else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    # own kernels enqueued per tick on this rank (graph replays re-run what was captured: count one eager tick)
    launches = torch.tensor([float(getattr(stage.engines[0], "launches_per_step", 0) or 0)], device=dev)
    dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    # ---- e2e: every tick the first stage takes the input token of the sequence it is about to run from pinned host
    # memory (H2D into that engine's token buffer, which the step consumes) and the last stage returns the emitted token
    # id to pinned host memory (D2H), with a stream sync per tick on every rank; wall clock, max over ranks
    h_tok = torch.randint(0, vocab, (1, 1), generator=g).pin_memory()
    dist.barrier()
    t0 = time.perf_counter()
    feed = (lambda slot: stage.engines[slot].tok.copy_(h_tok, non_blocking=True)) if stage.first else None
    for _ in range(a.steps):
        pipe.tick(pre_step=feed)
        if stage.last:
            h_tok.copy_(stage.engines[(pipe.t - 1 - rank) % S].tok, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    from .modeling import quantized_bytes
    code_bytes = torch.tensor([float(quantized_bytes(stage.model))], device=dev)
    dist.all_reduce(code_bytes, op=dist.ReduceOp.SUM)
    persistent = stage.engines[0].persistent is not None
    handoff = "NCCL p2p ring exchange per tick (host-issued)"
    if mailbox is not None:
        nerr = torch.tensor([mailbox.errors()], device=dev)
        dist.all_reduce(nerr, op=dist.ReduceOp.SUM)
        if nerr.item():
            raise RuntimeError(f"peer-memory hand-off: {int(nerr.item())} waits timed out")
        handoff = "device-side hand-off over NVLink peer memory (store + flag, no host sync, no NCCL on the tick path)"
        mailbox.close()
    out = {"handoff": handoff,"tok_s": a.steps / (ms * 1e-3), "ms_per_step": ms / a.steps, "e2e_tok_s": a.steps / t_e2e.item(),
           "launches": int(launches.item()) * a.steps, "clocks": ck, "code_bytes": code_bytes.item(),
           "engine": "persistent whole-step kernel per stage" if persistent else "grouped launches (6 per decoder layer)"}
    del stage, pipe
    torch.cuda.empty_cache()
    return out


def run_pipeline_bench(a, metric, clock_sampler_cls=None, peak_gbs=None, single_gpu_fn=None):
    """bench.py body for WORLD_SIZE > 1 (launched by torchrun, one rank per GPU).  Workload = BASELINE config 5: the
    Llama-2-70B layer pipeline; the 7B pipeline of round 1 is measured as well and reported under `secondary`."""
    import json
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    main = _pipeline_measure(a, a.model, world, rank, local, dev, clock_sampler_cls)
    second = None
    if a.model != "llama2-7b" and not getattr(a, "no_secondary", False):
        second = _pipeline_measure(a, "llama2-7b", world, rank, local, dev, None)
    # the N = 1 point of the SAME workload, measured in this run on rank 0's GPU (the other ranks wait): the whole model on
    # one device through the same engine -- the denominator of the scaling efficiency of this line
    single = None
    if single_gpu_fn is not None and not getattr(a, "no_single", False):
        torch.cuda.synchronize()
        if rank == 0:
            single = single_gpu_fn(a, torch, dev, peak_gbs, a.model)
        dist.barrier()
    if rank == 0:
        line = {
            "metric": metric, "value": main["tok_s"], "unit": "tokens/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 weights x int16 fixed-point activations (exact int32 dp4a), fp16 in/out",
            "data": "synthetic",
            "config": {"workload": f"{a.model} {a.codebook} bs=1 greedy decode, layer pipeline over {world} GPUs "
                                   f"({world} independent bs=1 sequences in flight, one per stage), random-init "
                                   f"packed weights, synthetic {a.prompt_len}-token prompts",
                       "parallelism": f"pp{world} (whole decoder layers per stage; {main['handoff']})",
                       "engine": main["engine"],
                       "l2": "inputs larger than L2 (each stage streams its slice of the %.1f GB of codes per tick)"
                             % (main["code_bytes"] / 1e9)},
            "e2e": {"value": main["e2e_tok_s"], "unit": "tokens/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8,
                    "how": "per tick: input token id pinned host -> the running engine's token buffer (first stage) and "
                           "emitted token id device -> pinned host (last stage), stream sync on every rank, wall clock, "
                           "max over ranks"},
            "gpu_launches": main["launches"], "clocks": main["clocks"],
        }
        if peak_gbs:
            # aggregate roofline: one token leaves per tick and every stage streams its slice once per tick
            roof = world * peak_gbs * 1e9 / main["code_bytes"]
            line["model_roofline"] = {"packed_code_bytes_per_token": main["code_bytes"],
                                      "tok_s_at_hbm_peak_all_gpus": roof, "frac_of_hbm_roofline": main["tok_s"] / roof}
        if single is not None:
            line["config"]["scaling_note"] = ("the N = 1 bench line is the metric's own workload (llama2-7b on one GPU); the N = 1 "
                                              "point of THIS workload is measured in this run: single_gpu_same_workload / "
                                              "scaling_vs_single_gpu")
            line["single_gpu_same_workload"] = single
            if "value" in single:
                line["scaling_vs_single_gpu"] = {"speedup": main["tok_s"] / single["value"],
                                                 "efficiency": main["tok_s"] / single["value"] / world}
        if second is not None:
            line["secondary"] = {"workload": f"llama2-7b {a.codebook} bs=1 decode, same pipeline ({second['engine']})",
                                 "value": second["tok_s"], "unit": "tokens/s", "ms_per_step": second["ms_per_step"],
                                 "e2e": second["e2e_tok_s"]}
        print(json.dumps(line))
    dist.destroy_process_group()
