"""The C++ host layer (idefix_b200/host): a problem directory written against the reference's
user-facing API (definitions.hpp + setup.cpp + idefix.ini) is compiled with nvcc against our
headers, linked to libidefix_b200.so and run as ./idefix -- the same way a reference user builds
and runs a problem.

CPU part: every problem compiles and links here (no device needed).
GPU part: the dumps ./idefix writes after N cycles are bit-identical with the dumps the
reference's own CPU build wrote for the same ini (tests/golden/short), the VTK file has the
reference's layout, restart from a dump continues bit-identically, and user-defined boundaries
written as device loops (idefix_for + KOKKOS_LAMBDA) behave like the built-in ones.
"""
import os
import re
import shutil
import struct
import subprocess

import numpy as np
import pytest

from cases import HOST_SHORT, SHORT, ini_for
from idefix_b200.dump import readDump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "idefix_b200", "host")
PROBLEMS = {"sod": ("sod", "NO"), "ot2d": ("OrszagTang", "YES"), "ot3d": ("OrszagTang3D", "YES"),
            "sodUserBC": ("sodUserBC", "NO"), "sph": ("SphericalMHD", "YES"), "polar": ("PolarMHD", "YES"),
            "sodiso": ("sodIso", "NO"), "iso25d": ("IsoMHD25D", "YES"), "sph25d": ("Spherical25D", "YES"),
            "polarhd": ("PolarHD2D", "NO"), "sodstrong": ("sodStrong", "NO"), "othooks": ("otHooks", "YES"),
            "axis": ("AxisMHD", "YES"), "axis25d": ("Axis25D", "YES")}


def build(problem, order, tmp_path_factory, vp=False):
    pdir, mhd = PROBLEMS[problem]
    out = tmp_path_factory.mktemp(f"{problem}_o{order}{'_vp' if vp else ''}") / "idefix"
    cmd = ["make", "-s", "-C", HOST, f"PROBLEM={os.path.join(HOST, 'problems', pdir)}", f"MHD={mhd}",
           f"ORDER={order}", f"OUT={out}"] + (["VP=1"] if vp else [])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return str(out)


_built = {}


def binary(problem, order, tmp_path_factory, vp=False):
    key = (problem, order, vp)
    if key not in _built:
        _built[key] = build(problem, order, tmp_path_factory, vp)
    return _built[key]


def run_idefix(exe, workdir, ini_text, *args):
    os.makedirs(workdir, exist_ok=True)
    with open(os.path.join(workdir, "idefix.ini"), "w") as fh:
        fh.write(ini_text)
    res = subprocess.run([exe, "-i", "idefix.ini", *args], cwd=workdir, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


# (a subset keeps the CPU suite short; every problem is built again by the GPU tests below)
@pytest.mark.parametrize("problem", ["sodUserBC", "ot3d", "sph", "iso25d", "axis"])
def test_problem_compiles_against_host_layer(problem, tmp_path_factory):
    exe = binary(problem, 2, tmp_path_factory)
    assert os.access(exe, os.X_OK)


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="reference tree not present on this box")
@pytest.mark.parametrize("ref_problem,mhd", [("HD/sod", "NO"), ("MHD/OrszagTang3D", "YES")])
def test_unmodified_reference_setups_compile(ref_problem, mhd, tmp_path):
    """drop-in check: the reference's own definitions.hpp + setup.cpp build against our headers"""
    out = tmp_path / "idefix"
    cmd = ["make", "-s", "-C", HOST, f"PROBLEM=/root/reference/test/{ref_problem}", f"MHD={mhd}", f"OUT={out}"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and os.access(out, os.X_OK), res.stdout[-3000:] + res.stderr[-3000:]


HOST_CASES = [c for c in SHORT if c.name in ("sod_hllc_plm", "sod_hllc_limo3", "ot2d_hlld_ucthlld_plm",
                                             "ot2d_hlld_contact_plm", "ot3d_hlld_contact_plm",
                                             "ot3d_hlld_ucthlld_plm", "ot3d_hlld_contact_plm_tr2")] + \
    [c for c in HOST_SHORT if not c.ref.endswith("_vp")]


def assert_same_fields(mine, gold):
    checked = 0
    for name, ref in gold.data.items():
        if name.startswith(("Vc-", "Vs-")):
            assert name in mine.data, name
            a, b = np.asarray(mine.data[name]), np.asarray(ref)
            if not np.array_equal(a, b):   # say where: the first differing index [i, j, k] and how many differ
                bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
                raise AssertionError(f"{name}: {len(bad)} of {a.size} values differ, max |diff| "
                                     f"{np.nanmax(np.abs(a - b)):.3e}, first at {bad[0].tolist()}, last at {bad[-1].tolist()}, "
                                     f"mine {a[tuple(bad[0])]!r} reference {b[tuple(bad[0])]!r}")
            checked += 1
    assert checked > 0
    for attr in ("x1", "x1l", "x1r", "x2", "x3"):
        assert np.array_equal(getattr(mine, attr), getattr(gold, attr))
    # every scalar the reference registers (time, dt, output bookkeeping, file numbers, geometry, periodicity):
    # a dump written here must restart the reference and vice versa (dump.cpp:628-810)
    scalars = [n for n in gold.data if not n.startswith(("Vc-", "Vs-"))]
    assert {"time", "dt", "dumpFileNumber", "dumpLast", "vtkFileNumber", "geometry", "periodicity"} <= set(scalars)
    for name in scalars:
        assert name in mine.data, name
        assert np.array_equal(np.asarray(mine.data[name]), np.asarray(gold.data[name])), name
    assert set(mine.data) == set(gold.data)


@pytest.mark.gpu
@pytest.mark.parametrize("case", HOST_CASES, ids=[c.name for c in HOST_CASES])
def test_cpp_host_dumps_equal_reference_dumps(case, tmp_path_factory, tmp_path):
    exe = binary(case.setup, case.order, tmp_path_factory)
    ncyc = max(case.cycles)
    out = run_idefix(exe, str(tmp_path), ini_for(case, dump_every_cycle=True), "-maxcycles", str(ncyc))
    assert "Main: Perfs are" in out and "cell updates/second" in out
    for n in (0,) + tuple(case.cycles):
        mine = readDump(os.path.join(tmp_path, f"dump.{n:04d}.dmp"))
        gold = readDump(os.path.join(ROOT, "tests", "golden", "short", f"{case.name}.c{n:04d}.dmp"))
        assert_same_fields(mine, gold)
        assert float(mine.data["time"][0]) == float(gold.data["time"][0])


VP_CASES = [c for c in SHORT + HOST_SHORT if c.ref.endswith("_vp")]


@pytest.mark.gpu
@pytest.mark.parametrize("case", VP_CASES, ids=[c.name for c in VP_CASES])
def test_cpp_host_vector_potential_dumps_equal_reference_dumps(case, tmp_path_factory, tmp_path):
    """EVOLVE_VECTOR_POTENTIAL through the C++ host (make VP=1): the problem set-ups carry the reference's own
    #ifdef EVOLVE_VECTOR_POTENTIAL branch (d.Ve(AX3e, ...)); dumps incl. the Ve-AX?e fields equal the dumps of a
    reference build with -DEVOLVE_VECTOR_POTENTIAL"""
    assert len(VP_CASES) >= 4
    exe = binary(case.setup, case.order, tmp_path_factory, vp=True)
    ncyc = max(case.cycles)
    run_idefix(exe, str(tmp_path), ini_for(case, dump_every_cycle=True), "-maxcycles", str(ncyc))
    for n in (0,) + tuple(case.cycles):
        mine = readDump(os.path.join(tmp_path, f"dump.{n:04d}.dmp"))
        gold = readDump(os.path.join(ROOT, "tests", "golden", "short", f"{case.name}.c{n:04d}.dmp"))
        assert any(k.startswith("Ve-") for k in gold.data)
        for name, ref in gold.data.items():
            if name.startswith(("Vc-", "Vs-", "Ve-")):
                assert name in mine.data, name
                assert np.array_equal(np.asarray(mine.data[name]), np.asarray(ref)), name


@pytest.mark.gpu
def test_cpp_host_batched_run_equals_per_cycle_run(tmp_path_factory, tmp_path):
    """-nowrite: cycles are enqueued back to back on the device (no host round trip between them);
    the run must reach the same time, printed with 17 significant digits, as the per-cycle run"""
    case = [c for c in SHORT if c.name == "ot3d_hlld_contact_plm"][0]
    exe = binary(case.setup, case.order, tmp_path_factory)
    batched = run_idefix(exe, str(tmp_path / "batched"), ini_for(case), "-maxcycles", "10", "-nowrite")
    percycle = run_idefix(exe, str(tmp_path / "percycle"), ini_for(case, dump_every_cycle=True), "-maxcycles", "10")
    ta = re.search(r"Main: Reached t=(\S+)", batched).group(1)
    tb = re.search(r"Main: Reached t=(\S+)", percycle).group(1)
    assert ta == tb and float(ta) > 0.0
    assert not os.path.exists(tmp_path / "batched" / "dump.0000.dmp")


@pytest.mark.gpu
def test_cpp_host_restart_is_bit_identical(tmp_path_factory, tmp_path):
    case = [c for c in SHORT if c.name == "ot2d_hlld_ucthlld_plm"][0]
    exe = binary(case.setup, case.order, tmp_path_factory)
    ini = ini_for(case, dump_every_cycle=True)
    run_idefix(exe, str(tmp_path), ini, "-maxcycles", "10")
    straight = readDump(str(tmp_path / "dump.0010.dmp"))
    again = tmp_path / "restart"
    os.makedirs(again)
    shutil.copy(tmp_path / "dump.0004.dmp", again / "dump.0004.dmp")
    out = run_idefix(exe, str(again), ini, "-restart", "4", "-maxcycles", "10")
    assert "Main: Restarting from dump file." in out
    assert_same_fields(readDump(str(again / "dump.0010.dmp")), straight)


@pytest.mark.gpu
def test_cpp_host_userdef_boundary_device_lambda(tmp_path_factory, tmp_path):
    """user-defined zero-gradient boundaries (idefix_for + KOKKOS_LAMBDA on hydro->Vc) == outflow"""
    case = [c for c in SHORT if c.name == "sod_hllc_plm"][0]
    exe = binary("sodUserBC", 2, tmp_path_factory)
    ini = ini_for(case, dump_every_cycle=True).replace("outflow", "userdef")
    run_idefix(exe, str(tmp_path), ini, "-maxcycles", "20")
    gold = readDump(os.path.join(ROOT, "tests", "golden", "short", "sod_hllc_plm.c0020.dmp"))
    assert_same_fields(readDump(str(tmp_path / "dump.0020.dmp")), gold)


@pytest.mark.gpu
def test_cpp_host_vtk_layout(tmp_path_factory, tmp_path):
    case = [c for c in SHORT if c.name == "ot2d_hlld_ucthlld_plm"][0]
    exe = binary(case.setup, case.order, tmp_path_factory)
    ini = ini_for(case).replace("[Output]", "[Output]\nvtk 1e-7")
    run_idefix(exe, str(tmp_path), ini, "-maxcycles", "2")
    raw = open(tmp_path / "data.0002.vtk", "rb").read()
    assert raw.startswith(b"# vtk DataFile Version 2.0\nIdefix ")
    assert b"\nBINARY\nDATASET RECTILINEAR_GRID\nFIELD FieldData 9\nGEOMETRY 1 1 int\n" in raw
    m = re.search(rb"\nDIMENSIONS (\d+) (\d+) (\d+)\n", raw)
    assert m and [int(v) for v in m.groups()] == [65, 65, 1]
    assert b"\nCELL_DATA 4096\n" in raw
    # scalars are big-endian floats in alphabetical order; RHO is uniform 25/(36 pi) at t=0 and stays close
    names = re.findall(rb"\nSCALARS (\w+) float\nLOOKUP_TABLE default\n", raw)
    assert names == [b"BX1", b"BX2", b"PRS", b"RHO", b"VX1", b"VX2"]
    off = raw.index(b"SCALARS RHO float\nLOOKUP_TABLE default\n") + len(b"SCALARS RHO float\nLOOKUP_TABLE default\n")
    rho = np.frombuffer(raw[off:off + 4 * 4096], dtype=">f4")
    dump = readDump(str(tmp_path / "dump.0000.dmp")) if os.path.exists(tmp_path / "dump.0000.dmp") else None
    assert np.all(np.isfinite(rho)) and abs(float(rho.mean()) - 25.0 / (36.0 * np.pi)) < 1e-3
    t = struct.unpack(">f", raw[raw.index(b"TIME 1 1 float\n") + 15:][:4])[0]
    assert t > 0.0


HOST_LOGIC_INI = """
# comment line
[Grid]
X1-grid    3  0.5  6  s+  1.0  8  u  3.0  5  l  9.0   # stretched | uniform | logarithmic patches
X2-grid    1  0.3  7  u  2.8
X3-grid    2  0.0  4  u  1.0  6  s-  4.0

[TimeIntegrator]
CFL         0.4
tstop       1.0
nstages     3

[Hydro]
solver    hlld      # trailing comment

[Boundary]
X1-beg    userdef
X1-end    outflow
X2-beg    reflective
X2-end    reflective
X3-beg    periodic
X3-end    periodic
"""


@pytest.mark.parametrize("with_axis", [False, True], ids=["reflective", "axis"])
def test_cpp_host_input_and_grid_match_python_mirror(tmp_path, with_axis):
    """CPU-only: the C++ Input / Grid (ini grammar, patch types u / l / s+ / s-, Newton solve of the stretch
    factor, GetOrSet's ostream round trip of defaults) produce bit-identical coordinates and the same values as
    the Python mirror, which tests/test_host.py pins against the reference's own grids.  with_axis: theta from 0 to
    pi on 16 cells between two axis boundaries -- the ghost spacing mirrored across the axis (gridHost.cpp:246-262),
    where rebuilding the faces outwards from the axis changes the outermost one by an ulp"""
    HOST_LOGIC_INI = globals()["HOST_LOGIC_INI"]
    if with_axis:
        HOST_LOGIC_INI = (HOST_LOGIC_INI.replace("X2-grid    1  0.3  7  u  2.8", "X2-grid    1  0.0  16  u  3.141592653589793")
                          .replace("X2-beg    reflective", "X2-beg    axis").replace("X2-end    reflective", "X2-end    axis"))
        assert "axis" in HOST_LOGIC_INI and "16  u" in HOST_LOGIC_INI
    from idefix_b200 import _abi
    from idefix_b200.grid import Definitions, Grid
    from idefix_b200.input import Input
    exe = tmp_path / "host_logic_check"
    src = os.path.join(ROOT, "tests", "host_cpp")
    cmd = ["nvcc", "-std=c++17", "-O1", "-x", "cu", "--extended-lambda", "--expt-relaxed-constexpr", "-fmad=false",
           "-gencode", "arch=compute_100a,code=sm_100a", "-DMHD=YES", "-DORDER=2", "-I", src, "-I", HOST,
           "-I", os.path.join(ROOT, "include"), os.path.join(src, "host_logic_check.cpp"), "-o", str(exe),
           "-L", os.path.join(ROOT, "idefix_b200"), "-lidefix_b200", "-Xlinker", "-rpath", "-Xlinker",
           os.path.join(ROOT, "idefix_b200"), "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    ini = tmp_path / "check.ini"
    ini.write_text(HOST_LOGIC_INI)
    out = subprocess.run([str(exe), "-i", str(ini), "-nolog"], capture_output=True, text=True, timeout=120, cwd=tmp_path)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    inp = Input.from_string(HOST_LOGIC_INI)
    g = Grid(inp, Definitions(3, 3, _abi.SPHERICAL, mhd=True, order=2))
    lines = out.stdout.splitlines()
    assert f"regular {int(g.isRegularCartesian)}" in lines
    got = {}
    for ln in lines:
        tok = ln.split()
        if tok[0] == "x":
            got[(int(tok[1]), int(tok[2]))] = [float.fromhex(v) for v in tok[3:7]]
        if tok[0] == "dir":
            d = int(tok[1])
            assert int(tok[3]) == g.np_int[d] and int(tok[5]) == g.np_tot[d] and int(tok[7]) == g.nghost[d]
            assert int(tok[9]) == g.lbound[d] and int(tok[11]) == g.rbound[d]
            assert float.fromhex(tok[13]) == g.xbeg[d] and float.fromhex(tok[15]) == g.xend[d]
    for d in range(3):
        for i in range(g.np_tot[d]):
            assert got[(d, i)] == [float(g.x[d][i]), float(g.xl[d][i]), float(g.xr[d][i]), float(g.dx[d][i])], (d, i)
    if with_axis:   # the mirrored ghosts differ from a plain continuation of the uniform grid (16 cells: by one ulp)
        n, ng = g.np_int[1], g.nghost[1]
        assert g.haveAxis and float(g.xr[1][ng + n + 1]) != 0.0 + (n + 2) * float(g.dx[1][ng])
    vals = {ln.split()[0]: ln.split()[1:] for ln in lines if ln.split()[0] in ("gamma", "cflmaxvar", "nstages", "solver", "check")}
    assert float.fromhex(vals["gamma"][0]) == inp.GetOrSet(float, "Hydro", "gamma", 0, 5.0 / 3.0) == 1.66667
    assert float.fromhex(vals["cflmaxvar"][0]) == 1.1 and vals["nstages"] == ["3"] and vals["solver"] == ["hlld"]
    assert vals["check"] == [str(inp.CheckEntry("Hydro", "emf")), str(inp.CheckEntry("Grid", "X1-grid")), "0"] == ["-1", "11", "0"]


@pytest.mark.gpu
def test_cpp_host_api_tour(tmp_path_factory, tmp_path):
    """problems/apiTour: user device array + idefix_for, idefix_reduce (Sum / Max), analysis function, user-defined
    VTK variables, create_mirror_view / deep_copy and a step hook (which sends TimeIntegrator::Cycle through the
    granular, reference-ordered ABI path)"""
    PROBLEMS["apitour"] = ("apiTour", "NO")
    exe = binary("apitour", 2, tmp_path_factory)
    ini = open(os.path.join(HOST, "problems", "apiTour", "idefix.ini")).read()
    out = run_idefix(exe, str(tmp_path), ini, "-maxcycles", "6")
    rows = [[float(v) for v in ln.split()] for ln in open(tmp_path / "analysis.txt")]
    assert len(rows) == 7                                  # t = 0 and one per cycle
    assert [int(r[4]) for r in rows] == list(range(7))     # the step hook ran once per cycle
    for t, mass, mass_host, vmax, _ in rows:
        assert abs(mass - mass_host) <= 1e-13 * mass_host  # device reduction == host sum
        assert abs(mass - rows[0][1]) <= 1e-13 * mass      # periodic box: mass is conserved to round-off
        assert 0.9 < vmax < 1.2
    assert rows[-1][0] > rows[0][0] == 0.0
    raw = open(tmp_path / "data.0000.vtk", "rb").read()
    names = re.findall(rb"\nSCALARS (\w+) float\nLOOKUP_TABLE default\n", raw)
    assert names == [b"BLOB", b"MACH", b"PRS", b"RHO", b"VX1", b"VX2"]
    off = raw.index(b"SCALARS BLOB float\nLOOKUP_TABLE default\n") + len(b"SCALARS BLOB float\nLOOKUP_TABLE default\n")
    blob = np.frombuffer(raw[off:off + 4 * 48 * 32], dtype=">f4")
    assert abs(float(blob.max()) - 0.5) < 0.02 and float(blob.min()) >= 0.0   # the user array made it into the file
