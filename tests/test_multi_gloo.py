"""N>1 host logic on CPU: world_size-2 (and 3) gloo runs of the shard → partial → all-gather →
final-add plumbing (ark_blst_b200/dist.py).  There is no GPU here, so each rank's partial comes
from the CPU oracle (the checker standing in for the device kernel); what is under test is the
partitioning, the collective and the combine order, against the unsharded oracle result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, g2, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ark_blst_b200 import dist as d
    from oracle import cref

    bases = cref.synth_bases(g2, 77, n, nthreads=2)
    scal = cref.synth_scalars(78, n, True)
    lo, hi = d.shard_range(n, rank, world)
    part = cref.msm(g2, bases[lo:hi], scal[lo:hi], 1, nthreads=2) if hi > lo else np.zeros(36 if g2 else 18, dtype=np.uint64)
    gathered = d.gather_partials(torch.from_numpy(part.view(np.int64).copy()), world)
    assert gathered.shape == (world, 36 if g2 else 18)
    assert np.array_equal(gathered[rank].numpy().view(np.uint64), part)
    if rank == 0:
        total = np.zeros_like(part)
        for r in range(world):
            total = cref.add(g2, total, gathered[r].numpy().view(np.uint64))
        whole = cref.msm(g2, bases, scal, 1, nthreads=2)
        ret.put(bool(cref.affine_equal(g2, total, whole)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,g2,n", [(2, 0, 1000), (2, 1, 301), (3, 0, 10), (2, 0, 1)])
def test_sharded_msm_gloo(world, g2, n):
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, g2, n, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get() is True


def test_shard_ranges_cover_exactly():
    from ark_blst_b200.dist import shard_range

    for n in (0, 1, 7, 1 << 20, (1 << 24) + 5):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1
