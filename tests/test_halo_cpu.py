"""world_size-2 gloo tests of the halo-exchange host logic (neighbour map, message ordering)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from idefix_b200.halo import coords_of, exchange_buffers, exchange_plan, neighbours, rank_of


def test_rank_layout_matches_cartesian_communicator():
    nproc = (2, 2, 2)
    for r in range(8):
        assert rank_of(coords_of(r, nproc), nproc) == r
    assert neighbours((0, 0, 0), nproc, [True] * 3, 0) == (1, 1)          # two blocks, periodic: same peer
    assert neighbours((1, 0, 1), nproc, [True] * 3, 2) == (1, 1)
    assert neighbours((0, 1, 0), (2, 2, 1), [False, False, True], 1) == (0, None)
    assert neighbours((2, 0, 0), (4, 1, 1), [True, True, True], 0) == (1, 3)
    assert neighbours((3, 0, 0), (4, 1, 1), [True, True, True], 0) == (2, 0)


def test_plan_order_same_peer():
    plan = exchange_plan(1, 1)
    assert plan == [("send", 0, 1), ("send", 1, 1), ("recv", 1, 1), ("recv", 0, 1)]
    assert exchange_plan(None, 3) == [("send", 1, 3), ("recv", 1, 3)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, periodic, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nproc = (world, 1, 1)
    left, right = neighbours(coords_of(rank, nproc), nproc, [periodic, True, True], 0)
    # send[side] carries (rank, side); after the exchange recv[0] must hold the left neighbour's
    # RIGHT slab and recv[1] the right neighbour's LEFT slab
    send = [torch.tensor([float(rank), 0.0]), torch.tensor([float(rank), 1.0])]
    recv = [torch.full((2,), -1.0), torch.full((2,), -1.0)]
    exchange_buffers(send, recv, exchange_plan(left, right))
    # dt reduction (MPI_Iallreduce MIN in the reference)
    dt = torch.tensor([1.0 + rank])
    dist.all_reduce(dt, op=dist.ReduceOp.MIN)
    out[rank] = (recv[0].tolist(), recv[1].tolist(), left, right, float(dt))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,periodic", [(2, True), (2, False), (3, True)])
def test_exchange_two_ranks_gloo(world, periodic):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, periodic, out), nprocs=world, join=True)
    for rank in range(world):
        r0, r1, left, right, dt = out[rank]
        assert dt == 1.0
        if left is not None:
            assert r0 == [float(left), 1.0], (rank, r0)
        else:
            assert r0 == [-1.0, -1.0]
        if right is not None:
            assert r1 == [float(right), 0.0], (rank, r1)
        else:
            assert r1 == [-1.0, -1.0]
