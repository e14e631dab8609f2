"""Helpers of the parity tests: run a case through the C ABI and compare with a dump."""
import numpy as np

from idefix_b200 import _abi
from idefix_b200.dump import vc_names
from idefix_b200.setups import REGISTRY
from idefix_b200.timeintegrator import run

# north-star tolerance: <= 1e-12 relative L1 on every evolved variable (BASELINE.json)
REL_L1_TOL = 1e-12


def wants_vector_potential(case):
    """EVOLVE_VECTOR_POTENTIAL cases: SHORT cases name a reference binary built with the define (`*_vp`), FULL cases
    a set-up `*_vp`"""
    return getattr(case, "ref", "").endswith("_vp") or case.setup.endswith("_vp")


def make(case, **kw):
    name = case.setup[:-3] if case.setup.endswith("_vp") else case.setup
    return REGISTRY[name](order=case.order, overrides=case.overrides, vector_potential=wants_vector_potential(case),
                          **kw).initialise()


def fields_of(setup):
    """interior Vc / Vs of a set-up's block, keyed and indexed like a dump ([i, j, k])"""
    d = setup.data
    Vc = d.download(_abi.FIELD_VC)
    kb, ke, jb, je, ib, ie = d.beg[2], d.end[2], d.beg[1], d.end[1], d.beg[0], d.end[0]
    out = {}
    for n, name in enumerate(vc_names(setup.definitions, setup.data.hydro.nTracer)):
        if name is not None:
            out[f"Vc-{name}"] = Vc[n, kb:ke, jb:je, ib:ie].transpose(2, 1, 0)
    if setup.definitions.mhd:
        Vs = d.download(_abi.FIELD_VS)
        for dd in range(setup.definitions.dimensions):
            o = [1 if e == dd else 0 for e in range(3)]
            out[f"Vs-BX{dd + 1}s"] = Vs[dd, kb:ke + o[2], jb:je + o[1], ib:ie + o[0]].transpose(2, 1, 0)
        if getattr(setup.definitions, "evolve_vector_potential", False):
            # edge-centred: one more point in the active directions perpendicular to the component (dump.cpp:921-930)
            Ve = d.download(_abi.FIELD_VE)
            nd = setup.definitions.dimensions
            comps = [(0, "AX1e", 0), (1, "AX2e", 1), (2, "AX3e", 2)] if nd == 3 else [(0, "AX3e", 2)]
            for n, name, dirn in comps:
                o = [1 if (e != dirn and e < nd) else 0 for e in range(3)]
                out[f"Ve-{name}"] = Ve[n, kb:ke + o[2], jb:je + o[1], ib:ie + o[0]].transpose(2, 1, 0)
    return out


def rel_l1(a, b):
    """relative L1 difference; falls back to absolute when the reference field is ~0"""
    den = np.abs(b).sum()
    num = np.abs(a - b).sum()
    scale = max(den, 1e-30 * a.size)
    return num / scale if den > 1e-12 * a.size else num / a.size


def compare_with_dump(setup, dump):
    """-> dict field -> (relative L1, max abs) for every field of the dump the block evolves"""
    mine = fields_of(setup)
    out = {}
    for name, ref in dump.data.items():
        if name in mine:
            a, b = mine[name], np.asarray(ref)
            assert a.shape == b.shape, (name, a.shape, b.shape)
            out[name] = (float(rel_l1(a, b)), float(np.abs(a - b).max()))
    assert out, "no common field"
    return out


def reference_metric(setup, dump):
    """pytools/idfx_test.py:414-428: mean over common 3-D fields of the RMS difference"""
    mine = fields_of(setup)
    errs = [np.sqrt(np.mean((mine[n] - np.asarray(dump.data[n])) ** 2)) for n in dump.data if n in mine]
    return float(sum(errs) / len(errs))


def advance_cycles(setup, n, fused):
    if fused:
        _abi.check(_abi.lib().idfx_cycles(setup.data._handle, n))
        setup.data.sync()
    else:
        _abi.check(_abi.lib().idfx_set_fused(setup.data._handle, 0))
        for _ in range(n):
            setup.tint.Cycle(setup.data)


def run_to_tstop(setup, fused=True):
    return run(setup.data, setup.tint, setup.tstop, fused=fused)
