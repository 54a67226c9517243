"""CPU tests of the multi-GPU host logic (world_size 2, gloo): the halo-exchange
schedule of the slab decomposition pairs up across ranks and moves exactly the
layers that changed."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from casmcode_clexmonte_b200.slab import halo_ops


def test_halo_schedule_pairs_up():
    for world in (1, 2, 4, 8):
        for n2, halo, Sk in ((8, 1, 2), (8, 2, 4), (4, 1, 2), (6, 3, 6)):
            N2 = n2 * world
            for kg in [None] + list(range(Sk)):
                sends, recvs = {}, {}
                for r in range(world):
                    for kind, k, peer in halo_ops(r, world, n2, halo, Sk, N2, kg):
                        if kind == "send":
                            assert 0 <= k < n2
                            sends.setdefault((r, peer), []).append((r * n2 + k) % N2)
                        else:
                            assert k < 0 or k >= n2
                            recvs.setdefault((peer, r), []).append((r * n2 + k) % N2)
                # every message that is sent is received, as the same global layer, in order
                assert sends == recvs, (world, n2, halo, Sk, kg)
                if kg is not None:
                    for lst in sends.values():
                        assert all(g % Sk == kg for g in lst)


def _worker(rank, world, port, n2, halo, Sk, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N2 = n2 * world
    layer = 5
    # local array of layers [-halo, n2+halo): owned layers hold (global k, version)
    loc = torch.full((n2 + 2 * halo, layer), -1.0)
    for k in range(n2):
        loc[k + halo] = float(rank * n2 + k)
    ok = True

    def exchange(kg):
        reqs = []
        for kind, k, peer in halo_ops(rank, world, n2, halo, Sk, N2, kg):
            t = loc[k + halo]
            reqs.append(dist.P2POp(dist.isend if kind == "send" else dist.irecv, t, peer))
        if reqs:
            for w in dist.batch_isend_irecv(reqs):
                w.wait()

    exchange(None)
    for k in range(-halo, n2 + halo):
        ok &= bool((loc[k + halo] == float((rank * n2 + k) % N2)).all())
    # now emulate sweeps: group g bumps every owned layer with global k % Sk == g by 1000,
    # then exchanges; ghosts must always equal the owner's current value when next read
    version = {k: 0 for k in range(N2)}
    for sweep in range(3):
        for g in range(Sk):
            for k in range(n2):
                if (rank * n2 + k) % Sk == g:
                    loc[k + halo] += 1000.0
            for k in range(N2):
                if k % Sk == g:
                    version[k] += 1
            exchange(g)
            for k in range(-halo, n2 + halo):
                gk = (rank * n2 + k) % N2
                ok &= bool((loc[k + halo] == float(gk + 1000 * version[gk])).all())
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("n2,halo,Sk", [(4, 1, 2), (8, 2, 4)])
def test_halo_exchange_gloo_world2(n2, halo, Sk):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + 7 * halo
    mp.spawn(_worker, args=(world, port, n2, halo, Sk, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
