"""Ghost-zone exchange between the blocks of a decomposed domain: the replacement of
``Mpi::ExchangeX1/X2/X3`` (src/mpi.cpp:247-846) and of the dt ``MPI_Iallreduce``
(src/timeIntegrator.cpp:318,358) for one process per GPU.

The slabs are packed / unpacked by the library's kernels (idfx_halo_pack / idfx_halo_unpack, the
reference's buffer layout mpi.cpp:79-134); this module only moves the packed buffers between
ranks -- ``torch.distributed`` point-to-point over NCCL (NVLink/NVSwitch inside a box) -- and
keeps everything on the block's stream, so an RK cycle never synchronises with the host.
``mode="p2p"`` hands the transport to the library altogether: the slabs are stored straight into the
neighbours' receive buffers over NVLink peer memory by the library's own kernels (CUDA IPC mappings,
idfx_p2p_connect), message numbers in device mailboxes order producer and consumer, and the dt minimum
travels the same way -- no NCCL call, no Python callback and no host synchronisation inside a cycle.

Rank layout follows the reference's Cartesian communicator (src/grid.cpp:222-223): rank =
xproc[0] + nproc[0]*(xproc[1] + nproc[1]*xproc[2]); periodic directions wrap.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _abi
from ._abi import check, lib


def rank_of(xproc, nproc):
    return xproc[0] + nproc[0] * (xproc[1] + nproc[1] * xproc[2])


def coords_of(rank, nproc):
    return (rank % nproc[0], (rank // nproc[0]) % nproc[1], rank // (nproc[0] * nproc[1]))


def neighbours(xproc, nproc, periodic, d):
    """(left rank, right rank) in direction d; None on a physical (non-periodic) boundary"""
    out = []
    for step in (-1, +1):
        c = list(xproc)
        c[d] += step
        if c[d] < 0 or c[d] >= nproc[d]:
            if not periodic[d]:
                out.append(None)
                continue
            c[d] %= nproc[d]
        out.append(rank_of(c, nproc))
    return tuple(out)


def exchange_plan(left, right):
    """Ordered point-to-point operations of one direction: (kind, side, peer).

    Side 0 = left.  Our left interior slab (send side 0) fills the LEFT neighbour's RIGHT ghosts and
    vice versa.  When left == right (two blocks, periodic) both messages travel between the same
    pair of ranks and are matched in posting order, hence: sends 0 then 1, receives 1 then 0."""
    plan = []
    if left is not None:
        plan.append(("send", 0, left))
    if right is not None:
        plan.append(("send", 1, right))
    if right is not None:
        plan.append(("recv", 1, right))
    if left is not None:
        plan.append(("recv", 0, left))
    return plan


def exchange_buffers(send, recv, plan, group=None):
    """run a plan on torch tensors send[side] / recv[side] (any device; NCCL or gloo)"""
    ops = []
    for kind, side, peer in plan:
        if kind == "send":
            ops.append(dist.P2POp(dist.isend, send[side], peer, group))
        else:
            ops.append(dist.P2POp(dist.irecv, recv[side], peer, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()   # NCCL: orders the current stream after the transfer, no host block


class _DevPtr:
    """__cuda_array_interface__ view of a raw device pointer (n doubles)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def as_tensor(ptr, n, device):
    return torch.as_tensor(_DevPtr(ptr, n), device=device)


class HaloExchanger:
    """Enrols the halo-exchange and reduction hooks of one block (one rank, one GPU)."""

    def __init__(self, data, nproc, xproc, mode="nccl"):
        if mode not in ("nccl", "p2p"):
            raise _abi.IdefixError(f"halo mode '{mode}' is not available in this build")
        self.mode = mode
        self.data = data
        if mode == "p2p":
            self._connect_p2p(data, tuple(nproc), tuple(xproc))
            return
        self.nproc, self.xproc = tuple(nproc), tuple(xproc)
        self.device = torch.device("cuda", torch.cuda.current_device())
        h = data._handle
        # everything (pack kernels, NCCL transfers, sweeps) is ordered on torch's current stream
        check(lib().idfx_set_stream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        periodic = [data.mygrid.lbound[d] == _abi.BC_PERIODIC for d in range(3)]
        self.plans, self.send, self.recv = {}, {}, {}
        for d in range(data.defs.dimensions):
            if nproc[d] == 1:
                continue
            n = C.c_size_t()
            check(lib().idfx_halo_buffer_size(h, d, C.byref(n)))
            self.send[d], self.recv[d] = [], []
            for side in range(2):
                ps, pr = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
                check(lib().idfx_halo_buffers(h, d, side, C.byref(ps), C.byref(pr)))
                self.send[d].append(as_tensor(C.cast(ps, C.c_void_p).value, n.value, self.device))
                self.recv[d].append(as_tensor(C.cast(pr, C.c_void_p).value, n.value, self.device))
            left, right = neighbours(self.xproc, self.nproc, periodic, d)
            self.plans[d] = exchange_plan(left, right)
        self.bytes_sent = 0
        data.enroll_halo_exchange(self._exchange)
        self._reduce_cb = _abi.REDUCE_FN(self._reduce)
        check(lib().idfx_enroll_reduce(h, self._reduce_cb, None))

    def _connect_p2p(self, data, nproc, xproc):
        """Peer-to-peer transport: every rank publishes the CUDA IPC handles of its receive buffers and of its
        mailbox, opens its neighbours' and hands the pointers to the library (idfx_p2p_connect).  From then on
        the exchange and the dt reduction are the library's own kernels over NVLink peer memory -- no Python,
        no NCCL call and no host synchronisation inside a cycle, and the cycle replays as a CUDA graph."""
        self.nproc, self.xproc = nproc, xproc
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.bytes_sent = 0
        h = data._handle
        check(lib().idfx_set_stream(h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        rank, world = rank_of(xproc, nproc), nproc[0] * nproc[1] * nproc[2]
        periodic = [data.mygrid.lbound[d] == _abi.BC_PERIODIC for d in range(3)]
        mine = {"recv": {}, "ptr": {}}
        mb_ptr, mb_handle = C.c_void_p(), (C.c_char * 64)()
        check(lib().idfx_p2p_mailbox(h, C.byref(mb_ptr), mb_handle))
        mine["mailbox"] = bytes(mb_handle.raw)
        dirs = [d for d in range(data.defs.dimensions) if nproc[d] > 1]
        for d in dirs:
            for side in range(2):
                hd = (C.c_char * 64)()
                check(lib().idfx_halo_ipc_handle(h, d, side, hd))
                mine["recv"][(d, side)] = bytes(hd.raw)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        opened = {}

        def open_handle(raw):
            if raw not in opened:
                p = C.POINTER(C.c_double)()
                check(lib().idfx_halo_open_ipc(h, C.c_char_p(raw), C.byref(p)))
                opened[raw] = C.cast(p, C.c_void_p).value
            return opened[raw]

        def mailbox_of(r):
            return mb_ptr.value if r == rank else open_handle(everyone[r]["mailbox"])

        peer_recv = ((C.c_void_p * 2) * 3)()
        peer_mb = ((C.c_void_p * 2) * 3)()
        for d in dirs:
            nb = neighbours(xproc, nproc, periodic, d)
            for side in range(2):
                if nb[side] is None:
                    continue
                # our side-`side` slab fills the ghosts of the OTHER side of the neighbour on `side`
                peer_recv[d][side] = open_handle(everyone[nb[side]]["recv"][(d, 1 - side)])
                peer_mb[d][side] = mailbox_of(nb[side])
        all_mb = (C.c_void_p * world)(*[mailbox_of(r) for r in range(world)])
        check(lib().idfx_p2p_connect(h, rank, world, peer_recv, peer_mb, all_mb))
        torch.cuda.synchronize()
        dist.barrier()   # nobody starts exchanging before every mailbox is mapped

    def _exchange(self, data, d):
        h = data._handle
        check(lib().idfx_halo_pack(h, d))
        exchange_buffers(self.send[d], self.recv[d], self.plans[d])
        check(lib().idfx_halo_unpack(h, d))
        self.bytes_sent += sum(self.send[d][s].numel() * 8 for k, s, p in self.plans[d] if k == "send")

    def _reduce(self, handle, dev_value, op, user):
        t = as_tensor(dev_value, 1, self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN if op == _abi.REDUCE_MIN else dist.ReduceOp.MAX)

    def close(self):
        if self.mode == "p2p":
            return
        self.send.clear()
        self.recv.clear()
