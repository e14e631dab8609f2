"""HD Roe solver on the GPU (riemann_roe_hd.cuh dispatched by the pencil sweeps).

FIRST-RUN TESTS.  The wiring was written after the round's GPU budget was spent: the solver function itself is
pinned on the CPU (host-compiled source vs oracle, tests/test_riemann_host.py; oracle vs the reference build and
the shipped HD/sod Roe goldens, tests/test_oracle.py), but these kernels have not run on a device yet.  Every
test here is therefore marked xfail(strict=False): a failure cannot mask the verified suite, a pass shows up as
XPASS.  The file sorts last so that nothing else runs after it in the same process.  Remove the marker once green.
"""
import os

import numpy as np
import pytest

import parity
from cases import ORACLE_FULL, ORACLE_SHORT
from idefix_b200 import _abi
from idefix_b200.dump import readDump
from idefix_b200.setups import REGISTRY
from oracle.godunov import Oracle
from test_oracle_gpu import build

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="HD Roe wiring has not run on a GPU yet (first-run tests)", strict=False)]

SHORT_ROE = [c for c in ORACLE_SHORT if c.setup in REGISTRY]


@pytest.mark.parametrize("case", SHORT_ROE, ids=[c.name for c in SHORT_ROE])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "granular"])
def test_roe_hd_short(case, fused):
    s = parity.make(case)
    done = 0
    for n in case.cycles:
        parity.advance_cycles(s, n - done, fused)
        done = n
        res = parity.compare_with_dump(s, readDump(os.path.join(GOLD, "short", f"{case.name}.c{n:04d}.dmp")))
        bad = {k: v for k, v in res.items() if not v[0] <= parity.REL_L1_TOL}
        assert not bad, f"after {n} cycles: {bad}"
    assert s.data.math_exceptions() == 0
    s.data.close()


@pytest.mark.parametrize("case", ORACLE_FULL, ids=[c.name for c in ORACLE_FULL])
def test_roe_hd_full(case):
    s = parity.make(case)
    parity.run_to_tstop(s, fused=True)
    gold = readDump(os.path.join(GOLD, "reference", case.golden))
    assert s.data.t == float(gold.data["time"][0])
    assert parity.reference_metric(s, gold) <= 1e-14
    assert s.data.math_exceptions() == 0
    s.data.close()


RANDOM = [
    # ndim ncomp n           nstages bc         seed
    (2, 2, (19, 7), 2, "periodic", 31),
    (2, 3, (8, 13), 3, "outflow", 32),
    (3, 3, (9, 6, 5), 2, "periodic", 33),
]


@pytest.mark.parametrize("case", RANDOM, ids=[f"{c[0]}d{c[1]}c_rk{c[3]}_{c[4]}" for c in RANDOM])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "granular"])
def test_roe_hd_cuda_equals_oracle_bitwise(case, fused):
    ndim, ncomp, n, nstages, bc, seed = case
    s = build(ndim, ncomp, False, n, "roe", None, nstages, bc, seed)
    s.InitFlow(s.data)
    o = Oracle.from_setup(s)
    o.set_state(s.data.host_mirror.Vc, None)
    s.data.SetBoundaries()
    parity.advance_cycles(s, 4, fused)
    o.cycle(4)
    assert s.data.t == o.t and s.data.dt == o.dt
    d = s.data
    kb, ke, jb, je, ib, ie = d.beg[2], d.end[2], d.beg[1], d.end[1], d.beg[0], d.end[0]
    Vc = d.download(_abi.FIELD_VC)
    assert np.array_equal(Vc[:, kb:ke, jb:je, ib:ie], o.Vc[:, kb:ke, jb:je, ib:ie])
    assert d.CheckNan() == 0 and d.math_exceptions() == 0
    o.close()
    d.close()
