"""NCCL, one process per GPU (needs >= 2 GPUs on the box; skipped otherwise): sharded pool scan with
the prototype all-gather and the ordered leaderboard hand-off gives boards bit-identical to one GPU."""
import importlib
import os
import socket

import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, c, k, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    gd = importlib.import_module("menghini-neurips23-code_b200.dist")
    eng = importlib.import_module("menghini-neurips23-code_b200.engine")
    f, t = synth.pool(n, c, peaked=0.1)
    rank_all = torch.from_numpy(synth.path_ranks(n)).to(torch.int32).to(dev)
    bounds = gd.shard_bounds(n, world)
    mine = gd.class_shards(c, world)[rank]
    protos = gd.gather_prototypes(t[mine.start:mine.stop].half().to(dev), c)
    assert torch.equal(protos.cpu(), t.half())
    board = gd.sharded_pool_scan(f[bounds[rank]:bounds[rank + 1]].half().to(dev), protos, 100.0, k, n, rank_all,
                                 lambda st: eng.Leaderboard(c, k, dev, state=st))
    q.put((rank, board.result()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("n,c,k", [(20000, 45, 16), (8000, 100, 16), (16000, 18, 700)])  # the last: set-mode boards (GRIP-sized k)
def test_sharded_scan_nccl_matches_single_gpu(n, c, k):
    import torch.multiprocessing as mp

    eng = importlib.import_module("menghini-neurips23-code_b200.engine")
    world = min(torch.cuda.device_count(), 8)
    f, t = synth.pool(n, c, peaked=0.1)
    one = eng.Leaderboard(c, k, "cuda:0")
    one.scan(f.half().cuda(), t.half().cuda(), 100.0,
             rank=torch.from_numpy(synth.path_ranks(n)).to(torch.int32).cuda())
    want = one.result()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, c, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r, res in got:
        assert res == want, r
