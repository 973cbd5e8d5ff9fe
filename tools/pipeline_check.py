"""2+ GPU check of the device-side pipeline hand-off (csrc/handoff.cu, parallel.PeerMailbox) against the host-issued NCCL
ring exchange: same stage weights, same prompts -> the token streams leaving the last stage must be identical, tick by
tick, and no wait may time out.  Also prints the tick time of both modes.
Usage (GPU box): torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/pipeline_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quip_for_all_b200.parallel import LlamaStage, PeerMailbox, RingPipeline  # noqa: E402


def run(mode, rank, world, dev, n_layers, ticks):
    S = world
    stage = LlamaStage("llama2-7b", "E8P12", rank, world, dev, S, cache_len=32 + 2 * ticks + 8 * S, num_hidden_layers=n_layers)
    mailbox = PeerMailbox(rank, world, S, dev) if mode == "peer" else None
    pipe = RingPipeline(stage, rank, world, S, mailbox=mailbox)
    g = torch.Generator().manual_seed(0)
    prompts = [torch.randint(0, 32000, (1, 32), generator=g) for _ in range(S)]
    stage.prefill_all(prompts, pipe)
    toks = []
    for _ in range(world - 1 + 4):
        pipe.tick()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ticks):
        emitted = pipe.tick()
        if emitted:
            toks.append(stage.engines[(pipe.t - 1 - rank) % S].tok.clone())
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / ticks], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    errs = mailbox.errors() if mailbox is not None else 0
    if mailbox is not None:
        mailbox.close()
    out = torch.cat(toks).flatten().cpu() if toks else torch.zeros(0, dtype=torch.long)
    del stage, pipe
    torch.cuda.empty_cache()
    return out, ms.item(), errs


def main():
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_layers, ticks = 4 * world, 48
    t_nccl, ms_nccl, _ = run("nccl", rank, world, dev, n_layers, ticks)
    t_peer, ms_peer, errs = run("peer", rank, world, dev, n_layers, ticks)
    ok = torch.tensor([1 if (torch.equal(t_nccl, t_peer) and errs == 0) else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == world - 1:
        print(f"tokens per mode: {t_nccl.numel()}  identical: {torch.equal(t_nccl, t_peer)}  wait timeouts: {errs}")
        print(f"tick: nccl {ms_nccl * 1e3:.1f} us   peer {ms_peer * 1e3:.1f} us   ({n_layers // world} 7B layers per stage)")
        print("first tokens:", t_peer[:8].tolist())
    dist.destroy_process_group()
    if ok.item() != 1:
        sys.exit(1)
    if rank == 0:
        print("PIPELINE_CHECK_OK")


if __name__ == "__main__":
    main()
