"""Multi-GPU plumbing for the pool scan and the prompt-tuning step: one process per GPU,
torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference does not shard this path at all — every DDP rank repeats the whole pool scan
(utils/clip_pseudolabels.py:55, methods/*/textual_fpl.py:214).  Here:
  * the pool is split into contiguous index ranges, rank r owns [bounds[r], bounds[r+1]);
  * each rank encodes ceil(C/G) class prompts and ONE all-gather assembles the [C,512] prototypes;
  * the leaderboard is order dependent, so its state (a few KB) travels rank 0 → 1 → … → G-1 while
    each rank replays only its own range (ordered hand-off); the last rank broadcasts the result.
All messages are tiny (≤ 100 KB): the design is latency-, not bandwidth-oriented.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int) -> List[int]:
    """Contiguous ranges of near-equal size; global order = rank order."""
    return [n * r // world for r in range(world + 1)]


def class_shards(c: int, world: int) -> List[range]:
    per = -(-c // world)
    return [range(min(r * per, c), min((r + 1) * per, c)) for r in range(world)]


def gather_prototypes(local: torch.Tensor, c: int, group=None) -> torch.Tensor:
    """local: this rank's [len(class_shards(c)[rank]), 512] prototypes → [c,512] on every rank with
    one all_gather (shards are padded to ceil(c/world) rows)."""
    world = dist.get_world_size(group)
    per = -(-c // world)
    pad = torch.zeros(per, local.shape[1], dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty(world * per, local.shape[1], dtype=local.dtype, device=local.device)
    if local.device.type == "cuda":
        dist.all_gather_into_tensor(out, pad, group=group)
    else:  # gloo (CPU tests)
        _all_gather_list(out, pad, world, group)
    return out[:c].contiguous()


def _all_gather_list(out, pad, world, group):
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out.copy_(torch.cat(parts))


def allreduce_mean_(t: torch.Tensor, group=None) -> torch.Tensor:
    """Prompt-gradient all-reduce (32 KiB for CoOp, 48 KiB for VPT): what DDP does for the only
    trainable tensor (methods/*/textual_prompt.py:131 via accelerator.backward)."""
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t /= dist.get_world_size(group)
    return t


def ordered_handoff(state: torch.Tensor, scan_own_range: Callable[[torch.Tensor], torch.Tensor],
                    group=None, ring: bool = False) -> torch.Tensor:
    """Runs `scan_own_range(state) -> state` on rank 0, 1, …, G-1 in that order, passing the
    leaderboard state along.  `state` must be initialised (empty boards) on rank 0; other ranks'
    input is overwritten by what they receive.  With ring=True the last rank passes the state back
    to rank 0 (the next batch continues there); otherwise the final state is broadcast so every rank
    returns the same boards."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return scan_own_range(state)

    def peer(r):  # send / recv / broadcast address processes by GLOBAL rank, also inside a sub-group
        return r if group is None else dist.get_global_rank(group, r)

    if rank > 0:
        dist.recv(state, src=peer(rank - 1), group=group)
    state = scan_own_range(state)
    if rank < world - 1:
        dist.send(state, dst=peer(rank + 1), group=group)
    if ring:
        if rank == world - 1:
            dist.send(state, dst=peer(0), group=group)
        if rank == 0:
            dist.recv(state, src=peer(world - 1), group=group)
    else:
        dist.broadcast(state, src=peer(world - 1), group=group)
    return state


def sharded_pool_scan(features_local: torch.Tensor, protos: torch.Tensor, scale: float, k: int,
                      n_total: int, rank_all: torch.Tensor, make_board, mode: int = 0, group=None,
                      timings: dict = None):
    """Exact pseudolabel boards for a pool sharded over the ranks.  features_local are this rank's
    rows of the pool (range shard_bounds(n_total, world)[rank]…); rank_all the global tie-break ranks.
    make_board(state_or_None) builds a Leaderboard (engine.Leaderboard on GPU).  Returns the board
    holding the final state (identical on every rank).

    Two phases, so that only the order-dependent part is serialised:
      1. every rank, concurrently: similarity + soft-max + arg-max over its own rows (the HBM pass),
         probabilities kept (`board.similarity`);
      2. rank 0 → 1 → … → G-1: conservative pre-filter against the bounds of the arriving state + exact replay
         of the surviving rows (`board.update`), then the few-KB state moves on.
    Per-row arithmetic is the one of the single-GPU fused scan (`board.scan`), so the boards are identical
    for every G.  `timings` (optional dict) receives CUDA-event handles / seconds of the two phases."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    bounds = shard_bounds(n_total, world)
    board = make_board(None)
    if world == 1:
        board.scan(features_local, protos, scale, mode=mode, idx0=bounds[rank], rank=rank_all)
        return board

    cuda = features_local.device.type == "cuda"
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if cuda and timings is not None else None
    if ev:
        ev[0].record()
    pred, _, probs = board.similarity(features_local, protos, scale, mode=mode)
    if ev:
        ev[1].record()

    def replay(state):
        b = make_board(state)
        if ev:
            ev[2].record()
        b.update(probs, pred, rank=rank_all, idx0=bounds[rank])
        if ev:
            ev[3].record()
        return b.state

    board = make_board(ordered_handoff(board.state, replay, group=group))
    if ev:
        torch.cuda.synchronize()
        timings["similarity_ms"] = ev[0].elapsed_time(ev[1])
        timings["replay_ms"] = ev[2].elapsed_time(ev[3])
    return board
