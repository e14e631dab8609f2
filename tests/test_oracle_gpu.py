"""GPU parity against the CPU oracle (oracle/godunov_oracle.c) on seeded random inputs, through the
C ABI: ragged and tiny grids, every solver / EMF scheme the oracle restates, RK1/2/3.
Bit-exact comparison: the CUDA path is built with -fmad=false (like the reference's own GPU
regression runs, pytools/idfx_test.py:119-124) and must reproduce the oracle to the last bit."""
import numpy as np
import pytest

import parity
from idefix_b200 import _abi
from idefix_b200.datablock import DataBlockHost
from idefix_b200.grid import Definitions
from idefix_b200.setups import Setup
from oracle.godunov import Oracle

pytestmark = pytest.mark.gpu


def make_ini(n, solver, emf, nstages, bc, mhd, ntracer=0):
    lines = ["[Grid]"]
    for d, nn in enumerate(n):
        lines.append(f"X{d + 1}-grid 1 0.0 {nn} u {1.0 + 0.5 * d}")
    lines += ["[TimeIntegrator]", "CFL 0.4", "tstop 100.0", "first_dt 1.e-4", f"nstages {nstages}",
              "[Hydro]", f"solver {solver}", "gamma 1.4"]
    if mhd and emf:
        lines.append(f"emf {emf}")
    if ntracer:
        lines.append(f"tracer {ntracer}")
    lines.append("[Boundary]")
    for d in range(len(n)):
        lines += [f"X{d + 1}-beg {bc}", f"X{d + 1}-end {bc}"]
    return "\n".join(lines) + "\n"


class RandomFlow(Setup):
    seed = 0

    def InitFlow(self, data):
        rng = np.random.default_rng(self.seed)
        d = DataBlockHost(data)
        h = data.hydro
        shape = d.Vc.shape[1:]
        d.Vc[h.RHO] = 1.0 + 0.3 * rng.random(shape)
        for c in range(data.defs.components):
            d.Vc[h.VX1 + c] = 0.4 * (rng.random(shape) - 0.5)
        d.Vc[h.PRS] = 1.0 + 0.3 * rng.random(shape)
        for n in range(h.nTracer):
            d.Vc[h.TRG + n] = rng.random(shape)
        if data.defs.mhd:
            nd = data.defs.dimensions
            for c in range(nd, data.defs.components):
                d.Vc[h.BX1 + c] = 0.2 * (rng.random(shape) - 0.5)
            if nd == 1:
                d.Vs[0][...] = 0.3
            else:
                dxmin = min(float(data.dx[a][0]) for a in range(nd))
                A = 0.3 * dxmin * rng.random((3,) + shape)
                d.MakeVsFromAmag(A)
        d.SyncToDevice()


def build(ndim, ncomp, mhd, n, solver, emf, nstages, bc, seed, ntracer=0, device=0):
    cls = type("RandomFlowCase", (RandomFlow,), {"defs": Definitions(ndim, ncomp, _abi.CARTESIAN, mhd=mhd), "seed": seed})
    return cls(order=2, ini=make_ini(n, solver, emf, nstages, bc, mhd, ntracer), device=device)


CASES = [
    # ndim ncomp mhd   n             solver  emf            nstages bc          seed
    (1, 1, False, (37,), "hllc", None, 2, "outflow", 1),
    (1, 3, False, (5,), "hll", None, 3, "periodic", 2),
    (2, 2, False, (19, 7), "tvdlf", None, 1, "periodic", 3),
    (3, 3, False, (9, 6, 5), "hllc", None, 2, "outflow", 4),
    (1, 2, True, (33,), "hlld", None, 2, "outflow", 5),
    (1, 3, True, (4,), "hll", None, 2, "periodic", 6),
    (2, 2, True, (23, 17), "hlld", "uct_hlld", 2, "periodic", 7),
    (2, 3, True, (8, 13), "hlld", "uct_contact", 3, "periodic", 8),
    (2, 2, True, (3, 2), "hll", "uct_hll", 2, "periodic", 9),          # fewer active cells than ghosts
    (3, 3, True, (13, 9, 11), "hlld", "uct_contact", 2, "periodic", 10),
    (3, 3, True, (7, 12, 5), "hlld", "uct_hlld", 3, "periodic", 11),
    (3, 3, True, (6, 5, 9), "hll", "arithmetic", 2, "periodic", 12),
    (3, 3, True, (5, 8, 6), "hlld", "uct0", 1, "periodic", 13),
    # passive tracers (10th entry = number of tracers)
    (3, 3, True, (9, 7, 8), "hlld", "uct_contact", 2, "periodic", 14, 2),
    (3, 3, True, (6, 9, 5), "hll", "uct_hll", 3, "periodic", 15, 1),
    (2, 2, False, (11, 6), "hllc", None, 2, "outflow", 16, 3),
]


@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}d{c[1]}c_{'mhd' if c[2] else 'hd'}_{c[4]}_{c[5]}_rk{c[6]}_{c[7]}" + (f"_tr{c[9]}" if len(c) > 9 else "") for c in CASES])
@pytest.mark.parametrize("fused", [True, False], ids=["fused", "granular"])
def test_cuda_equals_oracle_bitwise(case, fused):
    ndim, ncomp, mhd, n, solver, emf, nstages, bc, seed = case[:9]
    ntracer = case[9] if len(case) > 9 else 0
    s = build(ndim, ncomp, mhd, n, solver, emf, nstages, bc, seed, ntracer)
    s.InitFlow(s.data)
    host = s.data.host_mirror
    o = Oracle.from_setup(s)
    o.set_state(host.Vc, host.Vs)
    s.data.SetBoundaries()
    ncyc = 4
    parity.advance_cycles(s, ncyc, fused)
    o.cycle(ncyc)
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
        if ndim >= 2:
            assert d.hydro.CheckDivB() < 1e-12
    assert d.CheckNan() == 0
    assert d.math_exceptions() == 0   # every FP64 quotient / root was the correctly rounded one (fastmath.cuh)
    o.close()
    d.close()


def test_branch_free_division_and_sqrt_are_bit_identical(tmp_path):
    """tools/fastmath_check: the branch-free FP64 division / square root of the sweeps (csrc/fastmath.cuh)
    against nvcc's own operators on 16.7 M operand pairs -- random bit patterns (NaN, inf, subnormals),
    40 decades of magnitude, signed zeros: zero mismatches wherever nvcc's range test passes"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "check_fastmath"
    cmd = ["nvcc", "-arch=sm_100a", "-fmad=false", "-O3", "-I", os.path.join(root, "idefix_b200", "csrc"), "-o", str(exe),
           os.path.join(root, "tools", "fastmath_check", "check_fastmath.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert ", 0 mismatches," in out.stdout and "sqrt:" in out.stdout and out.stdout.count(" 0 mismatches") == 2, out.stdout


def test_enforce_boundary_dir_composes_to_set_boundaries():
    """idfx_enforce_boundary_dir (Boundary::EnforceBoundaryDir, boundary.hpp:287-385) direction by direction,
    then ReconstructVcField, fills the ghosts exactly like idfx_set_boundaries"""
    import ctypes as C
    outs = []
    for granular in (False, True):
        s = build(3, 3, True, (7, 6, 5), "hlld", "uct_contact", 2, "outflow", seed=21)
        s.InitFlow(s.data)
        h = s.data._handle
        if granular:
            for d in range(3):
                _abi.check(_abi.lib().idfx_enforce_boundary_dir(h, C.c_double(0.0), d))
            _abi.check(_abi.lib().idfx_reconstruct_vc_field(h, _abi.FIELD_VC))
        else:
            s.data.SetBoundaries()
        outs.append((s.data.download(_abi.FIELD_VC), s.data.download(_abi.FIELD_VS)))
        s.data.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


def test_two_blocks_on_one_gpu_peer_halo_bitwise():
    """Block decomposition on a SINGLE GPU: two blocks (2 x 1 x 1, periodic) in one process, each bound to the
    other's receive buffers with idfx_halo_bind_peer, so that idfx_halo_pack writes the slabs straight into the
    neighbour (the peer-to-peer path a multi-GPU run uses over NVLink).  The two blocks are advanced in lockstep
    through the granular entry points (the reference's TimeIntegrator::Cycle order, dt = min over blocks) and the
    assembled state must equal the dump the reference's own build wrote for the undecomposed run, bit for bit."""
    import ctypes as C
    import os

    from cases import SHORT
    from idefix_b200.dump import readDump
    from idefix_b200.setups import REGISTRY
    lib = _abi.lib()
    case = [c for c in SHORT if c.name == "ot3d_hlld_contact_plm_tr2"][0]
    nproc = (2, 1, 1)
    blocks = [REGISTRY[case.setup](order=case.order, overrides=case.overrides, nproc=nproc, xproc=(x, 0, 0)) for x in (0, 1)]
    hs = [s.data._handle for s in blocks]
    pd = C.POINTER(C.c_double)
    recv = []
    for h in hs:
        _abi.check(lib.idfx_set_stream(h, None))          # both blocks on the legacy default stream
        _abi.check(lib.idfx_set_fused(h, 0))
        sides = []
        for side in range(2):
            ps, pr = pd(), pd()
            _abi.check(lib.idfx_halo_buffers(h, 0, side, C.byref(ps), C.byref(pr)))
            sides.append(pr)
        recv.append(sides)
    for me, other in ((0, 1), (1, 0)):
        # our left slab fills the neighbour's RIGHT ghosts and vice versa (two blocks: both neighbours are `other`)
        _abi.check(lib.idfx_halo_bind_peer(hs[me], 0, 0, C.cast(recv[other][1], C.c_void_p)))
        _abi.check(lib.idfx_halo_bind_peer(hs[me], 0, 1, C.cast(recv[other][0], C.c_void_p)))

    def exchange_first(data, d):      # block 0: both blocks pack (peer stores), block 0 unpacks
        for h in hs:
            _abi.check(lib.idfx_halo_pack(h, d))
        _abi.check(lib.idfx_halo_unpack(hs[0], d))

    def exchange_second(data, d):     # block 1: its receive buffers were filled by block 0's pack
        _abi.check(lib.idfx_halo_unpack(hs[1], d))

    blocks[0].data.enroll_halo_exchange(exchange_first)
    blocks[1].data.enroll_halo_exchange(exchange_second)
    for s in blocks:
        s.InitFlow(s.data)
    datas = [s.data for s in blocks]
    for d in datas:
        d.SetBoundaries()
    tint = blocks[0].tint
    ncyc = max(case.cycles)
    for cyc in range(ncyc):
        t, dt = datas[0]._time()
        t0 = t
        for d in datas:
            d.ResetStage()
        newdt = None
        for stage in range(tint.nstages):
            for d in datas:
                d.SetBoundaries()
            for d in datas:
                d.PrimToCons()
                if tint.nstages > 1 and stage == 0:
                    d.state_copy()
                d.hydro.EvolveStage(t, dt)
            t += dt
            if stage == 0:
                newdt = tint.cfl * min(d.ComputeTimestep() for d in datas)
            if stage > 0:
                wc, w0 = tint.wc[stage - 1], tint.w0[stage - 1]
                for d in datas:
                    d.state_add_and_store(wc, w0)
                t = wc * t + w0 * t0
            for d in datas:
                d.set_time(t, dt)
                d.ConsToPrim()
        t = t0 + dt
        dt = tint.cflMaxVar * dt if newdt > tint.cflMaxVar * dt else newdt
        for d in datas:
            d.set_time(t, dt)
    gold = readDump(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "short", f"{case.name}.c{ncyc:04d}.dmp"))
    assert t == float(gold.data["time"][0])
    parts = [parity.fields_of(s) for s in blocks]
    checked = 0
    for name, ref in gold.data.items():
        if not name.startswith(("Vc-", "Vs-")):
            continue
        ref = np.asarray(ref)
        n0 = parts[0][name].shape[0]
        if name == "Vs-BX1s":       # the shared face belongs to both blocks (and must agree)
            assert np.array_equal(parts[0][name][-1], parts[1][name][0])
            full = np.concatenate([parts[0][name][:-1], parts[1][name]], axis=0)
        else:
            full = np.concatenate([parts[0][name], parts[1][name]], axis=0)
        assert full.shape == ref.shape, (name, full.shape, ref.shape, n0)
        assert np.array_equal(full, ref), name
        checked += 1
    assert checked >= 9     # 5 Vc fields + 2 tracers + 3 Vs components
    for d in datas:
        assert d.math_exceptions() == 0
        d.close()


def test_fast_division_envelope_violation_stops_the_run():
    """A state outside the validated envelope of the branch-free FP64 divisions (fastmath.cuh: every primitive zero or
    of magnitude 2^-140 .. 2^140) must not go unnoticed: the cycle's status carries bit 2 and the host layer raises
    (no silent loss of reference parity); idfx_math_exceptions counts the offending cells."""
    import ctypes as C
    from idefix_b200.timeintegrator import run
    s = build(2, 2, True, (16, 12), "hlld", "uct_contact", 2, "periodic", seed=77)
    s.InitFlow(s.data)
    host = s.data.host_mirror
    host.Vc[s.data.hydro.PRS, :, 5, 7] = 1.0e-200      # a near-vacuum cell: far below 2^-140
    host.SyncToDevice()
    s.data.SetBoundaries()
    t, dt, flags = C.c_double(), C.c_double(), C.c_int()
    _abi.check(_abi.lib().idfx_cycle(s.data._handle, C.byref(t), C.byref(dt), C.byref(flags)))
    assert flags.value & 4, flags.value
    assert s.data.math_exceptions() > 0
    with pytest.raises(RuntimeError, match="validated range"):
        run(s.data, s.tint, 1.0, fused=True, maxcycles=3)
    s.data.close()
