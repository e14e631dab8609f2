"""Peer-to-peer halo exchange (idfx_p2p_connect: the library's own send / receive kernels over peer memory, message
numbers in device mailboxes, dt minimum through the same mailboxes) -- block decomposition inside ONE process.

Every block of a decomposed SHORT case lives in this process (on cuda:0, or spread over the GPUs of the box when
there are several), each on its own stream.  The whole run is the fused path: idfx_cycles enqueues all cycles of a
block without blocking the host (CUDA-graph replay from the third cycle on, exchanges included), block after block,
and only then the host synchronises.  The assembled state must equal the dump the reference's own CPU build wrote
for the UNDECOMPOSED run, bit for bit (decomposition independence, the reference's MPI test strategy)."""
import ctypes as C
import os

import numpy as np
import pytest

import parity
from cases import SHORT
from idefix_b200 import _abi
from idefix_b200.dump import readDump
from idefix_b200.halo import coords_of, neighbours, rank_of
from idefix_b200.setups import REGISTRY

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def connect_in_process(blocks, nproc):
    """idfx_p2p_connect for blocks of one process: plain device pointers instead of IPC handles"""
    lib = _abi.lib()
    world = len(blocks)
    hs = [s.data._handle for s in blocks]
    ndim = blocks[0].definitions.dimensions
    periodic = [blocks[0].data.mygrid.lbound[d] == _abi.BC_PERIODIC for d in range(3)]
    mailbox, recv = [], []
    for h in hs:
        p = C.c_void_p()
        _abi.check(lib.idfx_p2p_mailbox(h, C.byref(p), None))
        mailbox.append(p.value)
        mine = {}
        for d in range(ndim):
            if nproc[d] == 1:
                continue
            for side in range(2):
                ps, pr = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
                _abi.check(lib.idfx_halo_buffers(h, d, side, C.byref(ps), C.byref(pr)))
                mine[(d, side)] = C.cast(pr, C.c_void_p).value
        recv.append(mine)
    devices = [s.data._desc.device for s in blocks]
    for a in set(devices):
        for b in set(devices):
            _abi.check(lib.idfx_p2p_enable_peer_access(a, b))
    for r, h in enumerate(hs):
        xproc = coords_of(r, nproc)
        peer_recv = ((C.c_void_p * 2) * 3)()
        peer_mb = ((C.c_void_p * 2) * 3)()
        for d in range(ndim):
            if nproc[d] == 1:
                continue
            nb = neighbours(xproc, nproc, periodic, d)
            for side in range(2):
                if nb[side] is not None:
                    peer_recv[d][side] = recv[nb[side]][(d, 1 - side)]
                    peer_mb[d][side] = mailbox[nb[side]]
        all_mb = (C.c_void_p * world)(*mailbox)
        _abi.check(lib.idfx_p2p_connect(h, r, world, peer_recv, peer_mb, all_mb))


def run_decomposed(case, nproc, ncyc):
    import torch
    lib = _abi.lib()
    world = nproc[0] * nproc[1] * nproc[2]
    ngpu = torch.cuda.device_count()
    # one host thread enqueues block after block: every kernel of this physics must already be loaded (CUDA loads
    # kernels lazily, and loading may wait for the device to drain -- which a receive kernel waiting for a block that
    # has not been enqueued yet would prevent), so an undecomposed block of the same case runs two cycles first
    for dev in range(min(ngpu, world)):
        warm = REGISTRY[case.setup](order=case.order, overrides=case.overrides, device=dev).initialise()
        parity.advance_cycles(warm, 5, True)
        warm.data.close()
    blocks = [REGISTRY[case.setup](order=case.order, overrides=case.overrides, nproc=nproc, xproc=coords_of(r, nproc),
                                   device=r % ngpu) for r in range(world)]
    assert [rank_of(coords_of(r, nproc), nproc) for r in range(world)] == list(range(world))
    connect_in_process(blocks, nproc)
    hs = [s.data._handle for s in blocks]
    for s in blocks:
        s.InitFlow(s.data)
    # one host thread drives every block: enqueue on ALL blocks (non-blocking calls), synchronise afterwards
    for h in hs:
        _abi.check(lib.idfx_set_boundaries(h, 0.0))
    for h in hs:
        _abi.check(lib.idfx_cycles(h, ncyc))
    for h in hs:
        _abi.check(lib.idfx_sync(h))
    return blocks


CASES = [("ot3d_hlld_contact_plm", (2, 1, 1)), ("ot3d_hlld_contact_plm", (2, 2, 2)), ("ot3d_hlld_ucthlld_plm", (1, 2, 2)),
         ("ot2d_hlld_ucthlld_plm", (2, 2, 1)), ("lw_fast_hlld_ppm_rk3", (1, 1, 2)), ("ot3d_hll_uct0_plm_rk3_tr2", (2, 1, 2))]


@pytest.mark.parametrize("name,nproc", CASES, ids=[f"{n}-{'x'.join(map(str, p))}" for n, p in CASES])
def test_p2p_blocks_in_one_process_equal_reference_dump(name, nproc):
    import torch
    # blocks sharing a GPU each need a hardware work queue of their own while a receive kernel waits for the
    # neighbour (CUDA_DEVICE_MAX_CONNECTIONS, 8 by default): at most four per device here
    if nproc[0] * nproc[1] * nproc[2] > 4 * torch.cuda.device_count():
        pytest.skip("more than four blocks per GPU")
    case = [c for c in SHORT if c.name == name][0]
    ncyc = max(case.cycles)
    blocks = run_decomposed(case, nproc, ncyc)
    gold = readDump(os.path.join(GOLD, "short", f"{name}.c{ncyc:04d}.dmp"))
    npi = [blocks[0].grid.np_int[a] // nproc[a] for a in range(3)]
    worst = 0.0
    for fname, ref in gold.data.items():
        if not fname.startswith(("Vc-", "Vs-")):
            continue
        full = np.full(np.asarray(ref).shape, np.nan)
        for r, s in enumerate(blocks):
            xp = coords_of(r, nproc)
            blk = parity.fields_of(s)[fname]
            lo = [xp[a] * npi[a] for a in range(3)]
            full[lo[0]:lo[0] + blk.shape[0], lo[1]:lo[1] + blk.shape[1], lo[2]:lo[2] + blk.shape[2]] = blk
        assert not np.isnan(full).any(), fname
        worst = max(worst, float(np.abs(full - np.asarray(ref)).max()))
    times = [s.data._time() for s in blocks]
    assert all(t == times[0] for t in times) and times[0][0] == float(gold.data["time"][0])
    for s in blocks:
        n = C.c_ulonglong()
        _abi.check(_abi.lib().idfx_p2p_timeouts(s.data._handle, C.byref(n)))
        assert n.value == 0 and s.data.math_exceptions() == 0
        # (no CheckDivB here: its maximum over the blocks is a collective, and one host thread cannot enter it for
        # every block at once; bitwise equality with the reference's dump covers the field)
    for s in blocks:
        s.data.close()
    assert worst == 0.0
