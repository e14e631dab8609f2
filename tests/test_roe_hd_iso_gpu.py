"""GPU parity of (1) the HD Roe solver (riemann_roe_hd.cuh, the solver of the reference's default test/HD/sod run),
including its 1-D strong-shock quirk (roeHD.hpp:312-328: the flux of the previous call stays in place -- sodstrong_*
cases, reproduced through SweepArgs::staleFlux) and (2) the isothermal kernels against the CPU oracle on random states.
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
from idefix_b200.datablock import DataBlockHost
from idefix_b200.grid import Definitions
from test_oracle_gpu import RandomFlow, build, make_ini

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
pytestmark = pytest.mark.gpu

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


class RandomIsoFlow(RandomFlow):
    def InitFlow(self, data):
        rng = np.random.default_rng(self.seed)
        d = DataBlockHost(data)
        h = data.hydro
        shape = d.Vc.shape[1:]
        d.Vc[h.RHO] = 1.0 + 0.3 * rng.random(shape)
        for c in range(data.defs.components):
            d.Vc[h.VX1 + c] = 0.4 * (rng.random(shape) - 0.5)
        if data.defs.mhd:
            nd = data.defs.dimensions
            for c in range(nd, data.defs.components):
                d.Vc[h.BX1 + c] = 0.2 * (rng.random(shape) - 0.5)
            dxmin = min(float(data.dx[a][0]) for a in range(nd))
            d.MakeVsFromAmag(0.3 * dxmin * rng.random((3,) + shape))
        d.SyncToDevice()


ISO_RANDOM = [
    # ndim ncomp mhd  n            solver  emf            nstages bc          seed
    (2, 2, False, (11, 6), "hllc", None, 2, "outflow", 41),
    (3, 3, False, (7, 5, 6), "tvdlf", None, 3, "periodic", 42),
    (2, 3, True, (8, 13), "hlld", "uct_contact", 2, "outflow", 43),
    (3, 3, True, (6, 9, 5), "roe", "uct_hll", 2, "periodic", 44),
    (2, 2, True, (9, 7), "hll", "uct0", 3, "periodic", 45),
    (3, 3, True, (5, 6, 7), "tvdlf", "arithmetic", 1, "outflow", 46),
]


@pytest.mark.parametrize("case", ISO_RANDOM, ids=[f"{c[0]}d{c[1]}c_{'mhd' if c[2] else 'hd'}_{c[4]}_{c[5]}_rk{c[6]}_{c[7]}" for c in ISO_RANDOM])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "granular"])
def test_isothermal_cuda_equals_oracle_bitwise(case, fused):
    ndim, ncomp, mhd, n, solver, emf, nstages, bc, seed = case
    ini = make_ini(n, solver, emf, nstages, bc, mhd).replace("gamma 1.4", "csiso constant 0.8")
    cls = type("RandomIsoCase", (RandomIsoFlow,),
               {"defs": Definitions(ndim, ncomp, _abi.CARTESIAN, mhd=mhd, isothermal=True), "seed": seed})
    s = cls(order=2, ini=ini, device=0)
    s.InitFlow(s.data)
    host = s.data.host_mirror
    o = Oracle.from_setup(s)
    o.set_state(host.Vc, host.Vs)
    s.data.SetBoundaries()
    parity.advance_cycles(s, 4, fused)
    o.cycle(4)
    assert s.data.t == o.t and s.data.dt == o.dt
    d = s.data
    kb, ke, jb, je, ib, ie = d.beg[2], d.end[2], d.beg[1], d.end[1], d.beg[0], d.end[0]
    Vc = d.download(_abi.FIELD_VC)
    assert np.array_equal(Vc[:, kb:ke, jb:je, ib:ie], o.Vc[:, kb:ke, jb:je, ib:ie])
    if mhd:
        Vs = d.download(_abi.FIELD_VS)
        for dd in range(ndim):
            off = [1 if e == dd else 0 for e in range(3)]
            sl = (dd, slice(kb, ke + off[2]), slice(jb, je + off[1]), slice(ib, ie + off[0]))
            assert np.array_equal(Vs[sl], o.Vs[sl])
        assert d.hydro.CheckDivB() < 1e-12
    assert d.CheckNan() == 0 and d.math_exceptions() == 0
    o.close()
    d.close()


HOST_ROE = [c for c in ORACLE_SHORT if c.setup not in REGISTRY]


@pytest.mark.parametrize("case", HOST_ROE, ids=[c.name for c in HOST_ROE])
def test_roe_hd_through_cpp_host(case, tmp_path_factory, tmp_path):
    """HD Roe through the C++ host layer (./idefix built from idefix_b200/host/problems/{sodIso,sodStrong}): the dumps
    equal those of the reference's own build for the same ini, bit for bit -- isothermal, and the strong blast tube
    whose interfaces take the 1-D stale-flux branch"""
    import test_host_cpp as thc
    from cases import ini_for
    assert len(HOST_ROE) >= 3
    exe = thc.binary(case.setup, case.order, tmp_path_factory)
    ncyc = max(case.cycles)
    thc.run_idefix(exe, str(tmp_path), ini_for(case, dump_every_cycle=True), "-maxcycles", str(ncyc))
    for n in (0,) + tuple(case.cycles):
        mine = readDump(os.path.join(tmp_path, f"dump.{n:04d}.dmp"))
        gold = readDump(os.path.join(GOLD, "short", f"{case.name}.c{n:04d}.dmp"))
        thc.assert_same_fields(mine, gold)
        assert float(mine.data["time"][0]) == float(gold.data["time"][0])
