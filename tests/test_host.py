"""CPU tests of the host-side mirror: ini grammar, grid generation, dump format, set-up initial
conditions (against the reference's own dump.0000), and the C-ABI surface."""
import os
import re
import subprocess

import numpy as np
import pytest

import parity
from cases import SHORT
from idefix_b200 import _abi
from idefix_b200.dump import readDump, vc_names, writeDump
from idefix_b200.grid import Definitions, Grid
from idefix_b200.input import Input
from idefix_b200.setups import REGISTRY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_input_grammar():
    inp = Input.from_string("""
# comment
[Grid]            # trailing comment
X1-grid    1  0.0  500  u  1.0
[Empty]
[Hydro]
solver  hllc   # the solver
flag    yes
""")
    assert inp.Get(int, "Grid", "X1-grid", 2) == 500
    assert inp.Get(float, "Grid", "X1-grid", 4) == 1.0
    assert inp.Get(str, "Hydro", "solver", 0) == "hllc"
    assert inp.Get(bool, "Hydro", "flag", 0) is True
    assert inp.CheckEntry("Hydro", "solver") == 1 and inp.CheckEntry("Hydro", "nope") == -1
    assert inp.CheckBlock("Empty") and not inp.CheckBlock("Nope")
    with pytest.raises(_abi.IdefixError):
        inp.Get(str, "Hydro", "missing", 0)
    with pytest.raises(_abi.IdefixError):
        Input.from_string("solver hllc\n")


def test_getorset_round_trips_through_six_digits():
    """src/input.hpp:147-170 stores the default through an ostream: gamma 5/3 -> 1.66667"""
    inp = Input.from_string("[Hydro]\nsolver hll\n")
    assert inp.GetOrSet(float, "Hydro", "gamma", 0, 5.0 / 3.0) == 1.66667
    assert inp.Get(float, "Hydro", "gamma", 0) == 1.66667
    assert inp.GetOrSet(float, "TimeIntegrator", "CFL_max_var", 0, 1.1) == 1.1


def _grid(text, defs):
    return Grid(Input.from_string(text), defs)


def test_uniform_grid_expressions():
    g = _grid("[Grid]\nX1-grid 1 0.0 10 u 1.0\n[Boundary]\nX1-beg outflow\nX1-end outflow\n", Definitions(1, 1))
    assert g.np_tot == [14, 1, 1] and g.nghost == [2, 0, 0] and g.isRegularCartesian
    dx = (1.0 - 0.0) / 10
    assert g.x[0][2] == 0.0 + 0.5 * dx and g.xl[0][0] == 0.0 + (-2) * dx and g.xr[0][13] == 0.0 + 12 * dx
    assert g.x[1][0] == 0.0 and g.dx[1][0] == 1.0          # unused dimension: [-0.5, 0.5]


def test_log_and_stretched_patches():
    txt = ("[Grid]\nX1-grid 3 1.0 8 l 2.0 16 u 4.0 8 s- 8.0\n"
           "[Boundary]\nX1-beg outflow\nX1-end outflow\n")
    g = _grid(txt, Definitions(1, 1))
    assert not g.isRegularCartesian and g.np_int[0] == 32
    xl, xr, dx = g.xl[0], g.xr[0], g.dx[0]
    assert np.allclose(xl[1:], xr[:-1], rtol=0, atol=1e-12)          # contiguous cells
    assert abs(xl[2] - 1.0) < 1e-14 and abs(xr[2 + 31] - 8.0) < 1e-9  # patch ends
    ratio = dx[3:9] / dx[2:8]
    assert np.allclose(ratio, ratio[0])                               # logarithmic: constant ratio
    assert np.all(np.diff(dx[2 + 24:2 + 32]) > 0)                     # stretched: growing cells
    # PPM needs 3 ghost cells (grid.cpp:61-65)
    g4 = _grid(txt, Definitions(1, 1, order=4))
    assert g4.nghost[0] == 3


def test_decomposition_boundaries():
    txt = ("[Grid]\nX1-grid 1 0.0 16 u 1.0\nX2-grid 1 0.0 8 u 1.0\n"
           "[Boundary]\nX1-beg periodic\nX1-end periodic\nX2-beg outflow\nX2-end reflective\n")
    g = Grid(Input.from_string(txt), Definitions(2, 2), nproc=(2, 2, 1), xproc=(1, 0, 0))
    sub = g.subdomain()
    assert sub["np_int"][:2] == [8, 4]
    assert sub["lbound"][0] == _abi.BC_INTERNAL and sub["rbound"][0] == _abi.BC_PERIODIC
    assert sub["lbound"][1] == _abi.BC_OUTFLOW and sub["rbound"][1] == _abi.BC_INTERNAL
    assert sub["x"][0][2] == g.x[0][2 + 8]
    with pytest.raises(_abi.IdefixError):
        Grid(Input.from_string(txt), Definitions(2, 2), nproc=(3, 1, 1))


def test_axis_boundaries_only_on_x2():
    """grid.cpp:98-104: `axis` is accepted on X2 (the library then insists on spherical geometry), refused elsewhere"""
    txt = ("[Grid]\nX1-grid 1 1.0 8 u 2.0\nX2-grid 1 0.0 8 u 3.141592653589793\nX3-grid 1 0.0 8 u 6.283185307179586\n"
           "[Boundary]\nX1-beg outflow\nX1-end outflow\nX2-beg axis\nX2-end axis\nX3-beg periodic\nX3-end periodic\n")
    sph = Definitions(3, 3, geometry=_abi.SPHERICAL, mhd=True)
    g = Grid(Input.from_string(txt), sph)
    assert g.haveAxis and g.lbound[1] == _abi.BC_AXIS and g.rbound[1] == _abi.BC_AXIS
    # gridHost.cpp:246-262: theta spacing mirrored across the axis, ghost faces rebuilt outwards from the axis face
    ng, n = g.nghost[1], g.np_int[1]
    assert g.xl[1][ng] == 0.0 and g.xr[1][ng - 1] == 0.0 and g.xl[1][ng - 1] == -g.dx[1][ng]
    assert g.xr[1][ng + n] == g.xr[1][ng + n - 1] + g.dx[1][ng + n - 1] and g.xl[1][ng + n] == g.xr[1][ng + n - 1]
    with pytest.raises(_abi.IdefixError):   # only on X2
        Grid(Input.from_string(txt.replace("X1-beg outflow", "X1-beg axis")), sph)
    with pytest.raises(_abi.IdefixError):   # only in spherical geometry
        Grid(Input.from_string(txt), Definitions(3, 3))
    with pytest.raises(_abi.IdefixError):   # the domain must reach the axis
        Grid(Input.from_string(txt.replace("8 u 3.141592653589793", "8 u 3.0")), sph)


@pytest.mark.parametrize("case", SHORT, ids=[c.name for c in SHORT])
def test_initial_conditions_match_reference_dump(case):
    """InitFlow of every set-up is bit-identical with the reference's dump.0000"""
    vp = parity.wants_vector_potential(case)
    s = REGISTRY[case.setup](order=case.order, overrides=case.overrides, device=None, vector_potential=vp)
    s.InitFlow(s.data)
    h, d = s.data.host_mirror, s.data
    ref = readDump(os.path.join(GOLD, "short", f"{case.name}.c0000.dmp"))
    kb, ke, jb, je, ib, ie = d.beg[2], d.end[2], d.beg[1], d.end[1], d.beg[0], d.end[0]
    for n, name in enumerate(vc_names(s.definitions)):
        if name is not None and not (vp and name.startswith("BX")):
            assert np.array_equal(h.Vc[n, kb:ke, jb:je, ib:ie].transpose(2, 1, 0), ref.data["Vc-" + name]), name
    if vp:
        # the set-up fills the edge-centred potential; B = curl A is derived on the device (tests/test_oracle_gpu.py)
        nd = s.definitions.dimensions
        comps = [(0, "AX1e", 0), (1, "AX2e", 1), (2, "AX3e", 2)] if nd == 3 else [(0, "AX3e", 2)]
        for n, name, dirn in comps:
            o = [1 if (e != dirn and e < nd) else 0 for e in range(3)]
            a = h.Ve[n, kb:ke + o[2], jb:je + o[1], ib:ie + o[0]].transpose(2, 1, 0)
            assert np.array_equal(a, ref.data["Ve-" + name]), name
    elif h.Vs is not None:
        for dd in range(s.definitions.dimensions):
            o = [1 if e == dd else 0 for e in range(3)]
            a = h.Vs[dd, kb:ke + o[2], jb:je + o[1], ib:ie + o[0]].transpose(2, 1, 0)
            assert np.array_equal(a, ref.data[f"Vs-BX{dd + 1}s"])
    for dd in range(3):
        g = s.grid.nghost[dd]
        assert np.array_equal(s.grid.x[dd][g:g + s.grid.np_int[dd]], getattr(ref, f"x{dd + 1}"))


def test_dump_round_trip(tmp_path):
    case = [c for c in SHORT if c.name == "ot2d_hlld_ucthlld_plm"][0]
    s = REGISTRY[case.setup](order=case.order, overrides=case.overrides, device=None)
    s.InitFlow(s.data)
    h = s.data.host_mirror
    path = tmp_path / "dump.0000.dmp"
    writeDump(str(path), s.data, vc_names(s.definitions), h.Vc, h.Vs, t=0.0, dt=1e-4)
    mine = readDump(str(path))
    ref = readDump(os.path.join(GOLD, "short", f"{case.name}.c0000.dmp"))
    for name, arr in ref.data.items():
        if name.startswith(("Vc-", "Vs-")):
            assert np.array_equal(mine.data[name], arr), name
    assert mine.data["dt"][0] == 1e-4 and int(mine.data["geometry"][0]) == 1
    # dump.N stores N+1 (dump.cpp:839), like the reference's own dump.0000
    assert int(mine.data["dumpFileNumber"][0]) == int(ref.data["dumpFileNumber"][0]) == 1
    assert [int(v) for v in mine.data["periodicity"]] == [1, 1, 0]
    # and the reference's own reader accepts the header
    assert mine.header.startswith("Idefix ") and "Dump Data" in mine.header


def test_host_only_block_has_no_device_path():
    s = REGISTRY["sod"](device=None)
    with pytest.raises(_abi.IdefixError):
        s.data.SetBoundaries()


def test_c_abi_exports_every_declared_symbol():
    """every function include/idefix_b200.h declares is exported by the library and bound"""
    header = open(os.path.join(ROOT, "include", "idefix_b200.h")).read()
    declared = set(re.findall(r"\b(idfx_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (idfx_[a-z0-9_]+)", out))
    assert declared <= exported, f"missing from the library: {sorted(declared - exported)}"
    assert declared == set(_abi.SYMBOLS), f"binding/header mismatch: {sorted(declared ^ set(_abi.SYMBOLS))}"
    _abi.lib()


def test_header_is_plain_c_and_matches_the_ctypes_mirror(tmp_path):
    """include/idefix_b200.h compiles as C11 with gcc (no C++ / CUDA / torch types at the boundary) and the
    layout of IdfxBlockDesc and the enum values equal what idefix_b200/_abi.py hands to the library"""
    import ctypes as C
    fields = [name for name, _ in _abi.IdfxBlockDesc._fields_]
    enums = {"IDFX_CARTESIAN": _abi.CARTESIAN, "IDFX_CYLINDRICAL": _abi.CYLINDRICAL, "IDFX_POLAR": _abi.POLAR,
             "IDFX_SPHERICAL": _abi.SPHERICAL, "IDFX_TVDLF": _abi.TVDLF, "IDFX_HLL": _abi.HLL, "IDFX_HLLC": _abi.HLLC,
             "IDFX_HLLD": _abi.HLLD, "IDFX_ROE": _abi.ROE, "IDFX_EMF_ARITHMETIC": _abi.EMF_ARITHMETIC,
             "IDFX_EMF_UCT0": _abi.EMF_UCT0, "IDFX_EMF_UCT_CONTACT": _abi.EMF_UCT_CONTACT,
             "IDFX_EMF_UCT_HLL": _abi.EMF_UCT_HLL, "IDFX_EMF_UCT_HLLD": _abi.EMF_UCT_HLLD,
             "IDFX_BC_INTERNAL": _abi.BC_INTERNAL, "IDFX_BC_PERIODIC": _abi.BC_PERIODIC,
             "IDFX_BC_REFLECTIVE": _abi.BC_REFLECTIVE, "IDFX_BC_OUTFLOW": _abi.BC_OUTFLOW,
             "IDFX_BC_SHEARINGBOX": _abi.BC_SHEARINGBOX, "IDFX_BC_AXIS": _abi.BC_AXIS,
             "IDFX_BC_USERDEF": _abi.BC_USERDEF, "IDFX_BC_UNDEFINED": _abi.BC_UNDEFINED,
             "IDFX_FIELD_VC": _abi.FIELD_VC, "IDFX_FIELD_UC": _abi.FIELD_UC, "IDFX_FIELD_VS": _abi.FIELD_VS,
             "IDFX_FIELD_INVDT": _abi.FIELD_INVDT, "IDFX_FIELD_EX": _abi.FIELD_EX, "IDFX_FIELD_EY": _abi.FIELD_EY,
             "IDFX_FIELD_EZ": _abi.FIELD_EZ, "IDFX_FIELD_UC0": _abi.FIELD_UC0, "IDFX_FIELD_VS0": _abi.FIELD_VS0,
             "IDFX_REDUCE_MIN": _abi.REDUCE_MIN, "IDFX_REDUCE_MAX": _abi.REDUCE_MAX}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "idefix_b200.h"', 'int main(void) {',
           '  printf("sizeof %zu\\n", sizeof(IdfxBlockDesc));']
    src += [f'  printf("{f} %zu %zu\\n", offsetof(IdfxBlockDesc, {f}), sizeof(((IdfxBlockDesc *)0)->{f}));' for f in fields]
    src += [f'  printf("{e} %d\\n", (int){e});' for e in enums]
    src += ['  return 0;', '}']
    cfile = tmp_path / "layout.c"
    cfile.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(cfile), "-o", str(exe)], check=True, capture_output=True, text=True)
    got = dict((ln.split()[0], [int(v) for v in ln.split()[1:]])
               for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert got.pop("sizeof") == [C.sizeof(_abi.IdfxBlockDesc)]
    for f in fields:
        fd = getattr(_abi.IdfxBlockDesc, f)
        assert got.pop(f) == [fd.offset, fd.size], f
    for e, v in enums.items():
        assert got.pop(e) == [v], e
    assert not got
    # no field of the header's struct is missing from the mirror
    body = re.search(r"typedef struct IdfxBlockDesc \{(.*?)\} IdfxBlockDesc;", open(os.path.join(ROOT, "include", "idefix_b200.h")).read(), re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = re.findall(r"\b\*?(\w+)(?:\[\d+\])?\s*;", body)
    assert declared == fields


def test_block_create_fails_loudly_without_gpu():
    import ctypes as C
    n = C.c_int(0)
    _abi.lib().idfx_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_abi.IdefixError, match="no CPU fallback"):
        REGISTRY["sod"]()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's own CPU build from oracle/_ref, no GPU involved) prints one
    JSON line with the keys the driver reads; skipped where the reference binary has not been built"""
    import json
    import subprocess
    import sys
    from oracle.ref_run import have
    if not have("ot3d_plm"):
        pytest.skip("oracle/_ref/ot3d_plm/idefix not built on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--cpu-block", "32",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "cell-updates/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"] == {"value": line["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("test/MHD/OrszagTang3D")
    # the arm describes what it ran (one OpenMP process on a bounded block), not the GPU arm's decomposition
    assert "decomposition" not in line["config"] and "halo" not in line["config"] and line["config"]["block"] == 32
