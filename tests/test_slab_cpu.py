"""CPU tests of the slab decomposition: ownership map properties, and the all-to-all plumbing
over a world_size-2 ``gloo`` group with the host mirror of the device layout math."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from __graft_entry__ import PKG_DIR, load_package


@pytest.fixture(scope="module")
def pkg():
    p = load_package()
    p.lib()
    return p


@pytest.mark.parametrize("ppd,G", [(16, 1), (16, 2), (32, 4), (64, 8), (2048, 8)])
def test_ownership_is_a_partition_with_hermitian_pairs_together(pkg, ppd, G):
    h = ppd // (2 * G)
    seen = {}
    for y in range(ppd):
        r, s = pkg.slab_owner(ppd, G, y)
        assert 0 <= r < G and 0 <= s < 2 * h
        assert (r, s) not in seen
        seen[(r, s)] = y
    assert len(seen) == ppd  # every slot of every rank is used exactly once
    for y in range(1, ppd // 2):
        assert pkg.slab_owner(ppd, G, y)[0] == pkg.slab_owner(ppd, G, ppd - y)[0]  # +ky and -ky rows on the same rank
    assert pkg.slab_owner(ppd, G, ppd // 2) == (0, h)  # the zero Nyquist row sits in rank 0's spare slot


@pytest.mark.parametrize("ppd,G", [(1024, 2), (1024, 8), (2048, 8)])
def test_ownership_balances_the_unmasked_modes(pkg, ppd, G):
    """Rows outside the k_cutoff sphere are masked (reference src/zeldovich.cpp:350-358), so the generation work of a
    primary row ky is ~ the area of the disc kx^2 + kz^2 < (ppd/2)^2 - ky^2; every rank must get the same share."""
    half = ppd // 2
    work = np.zeros(G)
    for y in range(half):
        r, _ = pkg.slab_owner(ppd, G, y)
        work[r] += np.pi * (half * half - y * y)
    assert work.max() / work.min() < 1.05, work / work.mean()  # blocks of rows: 2.2x (G=2) to 8.5x (G=8)


def _worker(rank, world, port, ppd, na, q):
    import importlib.util
    import sys

    sys.path.insert(0, os.path.dirname(PKG_DIR))
    from __graft_entry__ import load_package

    pkg = load_package()
    spec = importlib.util.spec_from_file_location("zplt_distributed", os.path.join(PKG_DIR, "distributed.py"))
    zd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(zd)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    N = ppd
    h = N // (2 * world)
    block = (N // world) * na * 2 * h * N
    send = torch.full((world * block,), -1.0, dtype=torch.float64)
    # every row this rank owns in stage 1 gets the label of its (a, z, y); x is the position inside the row
    label = lambda a, z, y: float((a * N + z) * N + y)
    for y in range(N):
        r, s = pkg.slab_owner(N, world, y)
        if r != rank:
            continue
        for a in range(na):
            for z in range(N):
                off = pkg.slab_offset(N, world, na, 1, rank, a, z, y)
                assert off >= 0 and off % N == 0
                send[off:off + N] = label(a, z, y) + torch.arange(N, dtype=torch.float64) / (2 * N)
    assert (send >= 0).all()  # the stage-1 buffer is exactly covered
    recv = torch.empty_like(send)
    zd.exchange_tensors(send, recv)
    ok = True
    z0, z1 = zd.plane_range(N, rank, world)
    for a in range(na):
        for z in range(z0, z1):
            for y in range(N):
                off = pkg.slab_offset(N, world, na, 2, rank, a, z, y)
                want = label(a, z, y) + torch.arange(N, dtype=torch.float64) / (2 * N)
                ok = ok and bool(torch.equal(recv[off:off + N], want))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_all_to_all_reassembles_planes_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 16, 2, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
