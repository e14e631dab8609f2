"""Multi-GPU parity: block decomposition over NCCL must reproduce the reference bit for bit."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nranks, case, nproc, mode="fused", port=29611, halo="nccl"):
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"),
           case, *[str(v) for v in nproc], mode, halo]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MGPU OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


@pytest.mark.parametrize("case,nproc", [("ot3d_hlld_contact_plm", (1, 2, 1)), ("ot3d_hlld_contact_plm_tr2", (2, 1, 1)), ("ot3d_hlld_ucthlld_plm", (2, 1, 1)),
                                        ("ot2d_hlld_ucthlld_plm", (1, 2, 1)), ("lw_fast_hlld_ppm_rk3", (1, 1, 2))])
def test_two_gpus(case, nproc):
    _run(2, case, nproc)


def test_two_gpus_granular():
    _run(2, "ot3d_hlld_contact_plm", (2, 1, 1), mode="granular")


@pytest.mark.parametrize("case,nproc", [("ot3d_hlld_contact_plm", (2, 1, 1)), ("ot3d_hlld_ucthlld_plm", (1, 1, 2)),
                                        ("ot2d_hlld_ucthlld_plm", (1, 2, 1)), ("lw_fast_hlld_ppm_rk3", (1, 2, 1))])
def test_two_gpus_p2p(case, nproc):
    """the library's own peer-to-peer exchange over NVLink (CUDA IPC, device mailboxes), cycle replayed as a graph"""
    _run(2, case, nproc, halo="p2p")


def test_two_gpus_p2p_granular():
    _run(2, "ot3d_hlld_contact_plm_tr2", (2, 1, 1), mode="granular", halo="p2p")


@pytest.mark.parametrize("case,nproc", [("ot3d_hlld_contact_plm", (2, 2, 2)), ("ot3d_hll_uct0_plm_rk3", (2, 2, 2))])
def test_eight_gpus(case, nproc):
    _run(8, case, nproc)


@pytest.mark.parametrize("case,nproc", [("ot3d_hlld_contact_plm", (2, 2, 2)), ("ot3d_hll_uct0_plm_rk3_tr2", (2, 2, 2))])
def test_eight_gpus_p2p(case, nproc):
    _run(8, case, nproc, halo="p2p")
