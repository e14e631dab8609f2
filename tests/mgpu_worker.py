"""torchrun worker of the multi-GPU parity test: runs a SHORT case block-decomposed over all ranks
(one rank per GPU, NCCL halo exchange) and checks on rank 0 that the assembled global state is
bit-identical with the dump the reference's CPU build wrote (decomposition independence, the
reference's own MPI test strategy: SURVEY.md section 4).

    torchrun --nproc-per-node N tests/mgpu_worker.py <case> <n1> <n2> <n3> [fused|granular] [nccl|p2p]

Transport: ``nccl`` = packed slabs through torch.distributed point-to-point; ``p2p`` = the library's own send /
receive kernels over NVLink peer memory (CUDA IPC mappings, device mailboxes, idfx_p2p_connect), no host callback,
the cycle replayed as a CUDA graph.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import parity  # noqa: E402
from cases import SHORT  # noqa: E402
from idefix_b200 import _abi  # noqa: E402
from idefix_b200.dump import readDump, vc_names  # noqa: E402
from idefix_b200.halo import HaloExchanger, coords_of  # noqa: E402
from idefix_b200.setups import REGISTRY  # noqa: E402


def main():
    name = sys.argv[1]
    nproc = tuple(int(v) for v in sys.argv[2:5])
    fused = not (len(sys.argv) > 5 and sys.argv[5] == "granular")
    halo = sys.argv[6] if len(sys.argv) > 6 else "nccl"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert nproc[0] * nproc[1] * nproc[2] == world
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    case = [c for c in SHORT if c.name == name][0]
    xproc = coords_of(rank, nproc)
    s = REGISTRY[case.setup](order=case.order, overrides=case.overrides, nproc=nproc, xproc=xproc, device=local)
    ex = HaloExchanger(s.data, nproc, xproc, mode=halo)
    s.InitFlow(s.data)
    dist.barrier()
    s.data.SetBoundaries()
    ncyc = max(case.cycles)
    if fused:
        parity.advance_cycles(s, ncyc, True)
    else:
        _abi.check(_abi.lib().idfx_set_fused(s.data._handle, 0))
        _abi.check(_abi.lib().idfx_cycles(s.data._handle, ncyc))
        s.data.sync()
    mine = parity.fields_of(s)
    divb = s.data.hydro.CheckDivB() if s.definitions.mhd else 0.0
    timeouts = 0
    if halo == "p2p":
        import ctypes as C
        n = C.c_ulonglong()
        _abi.check(_abi.lib().idfx_p2p_timeouts(s.data._handle, C.byref(n)))
        timeouts = n.value
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((xproc, {k: np.ascontiguousarray(v) for k, v in mine.items()}), gathered, dst=0)
    ok = True
    if rank == 0:
        gold = readDump(os.path.join(ROOT, "tests", "golden", "short", f"{name}.c{ncyc:04d}.dmp"))
        worst = 0.0
        for fname, ref in gold.data.items():
            if not fname.startswith(("Vc-", "Vs-")):
                continue
            full = np.full(ref.shape, np.nan)
            for xp, fields in gathered:
                blk = fields[fname]
                npi = [s.grid.np_int[a] // nproc[a] for a in range(3)]
                lo = [xp[a] * npi[a] for a in range(3)]
                full[lo[0]:lo[0] + blk.shape[0], lo[1]:lo[1] + blk.shape[1], lo[2]:lo[2] + blk.shape[2]] = blk
            worst = max(worst, float(np.abs(full - ref).max()))
        ok = worst == 0.0 and divb < 1e-12 and timeouts == 0
        print(f"MGPU {'OK' if ok else 'FAIL'} case={name} nproc={nproc} fused={fused} halo={halo} max|diff|={worst:.3e} "
              f"divB={divb:.2e} halo_bytes={ex.bytes_sent} p2p_timeouts={timeouts}", flush=True)
    ex.close()
    s.data.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
