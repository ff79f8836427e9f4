"""world_size-2 (and 3) gloo runs of the multi-GPU host logic on CPU: shard bounds, the prototype
all-gather, the prompt-gradient all-reduce and the ordered leaderboard hand-off.  The leaderboard
itself is played by the oracle's resumable Boards (the CUDA Leaderboard has the same
scan(features, protos, scale, mode, idx0, rank) / .state surface and is covered by -m gpu tests)."""
import importlib
import os
import pickle
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import leaderboard_ref, synth

STATE_BYTES = 1 << 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class PyBoard:
    """Oracle-backed stand-in with the Leaderboard surface; state = pickled boards in a uint8 tensor."""

    def __init__(self, c, k, state=None):
        self.c, self.k = c, k
        if state is None:
            self.boards = leaderboard_ref.Boards(c, k)
            self.state = torch.zeros(STATE_BYTES, dtype=torch.uint8)
            self._store()
        else:
            self.state = state
            n = int.from_bytes(bytes(state[:4].tolist()), "little")
            self.boards = pickle.loads(bytes(state[4:4 + n].tolist()))

    def _store(self):
        raw = pickle.dumps(self.boards)
        assert len(raw) + 4 <= STATE_BYTES
        buf = len(raw).to_bytes(4, "little") + raw
        self.state[:len(buf)] = torch.tensor(list(buf), dtype=torch.uint8)

    def scan(self, feats, protos, scale, mode=0, idx0=0, rank=None):
        _, probs, pred = leaderboard_ref.softmax_argmax(feats.numpy(), protos.numpy(), scale)
        self.boards.feed(probs, pred, rank.tolist(), idx0)
        self._store()

    def similarity(self, feats, protos, scale, mode=0):
        _, probs, pred = leaderboard_ref.softmax_argmax(feats.numpy(), protos.numpy(), scale)
        return pred, None, probs

    def update(self, probs, pred, rank=None, idx0=0):
        self.boards.feed(probs, pred, rank.tolist(), idx0)
        self._store()

    def result(self):
        return self.boards.result()


def _worker(rank, world, port, n, c, k, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gd = importlib.import_module("menghini-neurips23-code_b200.dist")
    f, t = synth.pool(n, c, peaked=0.1)
    rank_all = torch.from_numpy(synth.path_ranks(n))
    bounds = gd.shard_bounds(n, world)
    # prototypes: every rank contributes its class shard, one all-gather
    mine = gd.class_shards(c, world)[rank]
    protos = gd.gather_prototypes(t[mine.start:mine.stop], c)
    assert torch.equal(protos, t)
    # prompt-gradient all-reduce
    gsum = gd.allreduce_mean_(torch.full((4,), float(rank + 1)))
    assert torch.allclose(gsum, torch.full((4,), (world + 1) / 2))
    board = gd.sharded_pool_scan(f[bounds[rank]:bounds[rank + 1]], protos, 100.0, k, n, rank_all,
                                 lambda st: PyBoard(c, k, st))
    q.put((rank, board.result()))
    # ring hand-off over two consecutive "steps" (what bench.py does at N > 1)
    st = PyBoard(c, k).state
    for step in range(2):
        def scan(s, step=step):
            b = PyBoard(c, k, s)
            lo = (step * world + rank) * 10
            b.scan(f[lo:lo + 10], protos, 100.0, idx0=lo, rank=rank_all)
            return b.state
        st = gd.ordered_handoff(st, scan, ring=True)
    if rank == 0:
        q.put(("ring", PyBoard(c, k, st).result()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scan_matches_single_process(world):
    n, c, k = 600, 7, 5
    f, t = synth.pool(n, c, peaked=0.1)
    _, probs, pred = leaderboard_ref.softmax_argmax(f.numpy(), t.numpy(), 100.0)
    ranks = synth.path_ranks(n).tolist()
    want = leaderboard_ref.leaderboard(probs, pred, k, ranks)
    want_ring = leaderboard_ref.leaderboard(probs[:20 * world], pred[:20 * world], k, ranks)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, c, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world + 1)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for tag, res in got:
        if tag == "ring":
            assert res == want_ring
        else:
            assert res == want, tag


def test_shard_helpers():
    gd = importlib.import_module("menghini-neurips23-code_b200.dist")
    assert gd.shard_bounds(10, 4) == [0, 2, 5, 7, 10]
    assert gd.shard_bounds(0, 2) == [0, 0, 0]
    assert [list(r) for r in gd.class_shards(10, 8)] == [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9], [], [], []]
    assert [len(r) for r in gd.class_shards(102, 8)] == [13] * 7 + [11]
