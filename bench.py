#!/usr/bin/env python
"""bench.py -- cell-updates/s of the explicit Godunov (M)HD update (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c1..c5] [--block B] [--emf NAME]

Default workload (config.workload, BASELINE config C4): the reference's 3-D Orszag-Tang vortex
(test/MHD/OrszagTang3D, the set-up of its published benchmark, doc/source/performances.rst:5-10): HLLD + PLM +
constrained transport, RK2, periodic, FP64, B^3 interior cells PER GPU (default 512^3 = C4 on one GPU).  N > 1
(launched by torchrun, one rank per GPU) block-decomposes 2x1x1 / 2x2x1 / 2x2x2 with the halo exchange that
replaces Mpi::ExchangeX1/X2/X3 -- weak scaling, per-GPU block fixed; every N > 1 line also carries
`c4_as_stated` = the same measurement at 256^3 per GPU (at N = 8 exactly C4: 512^3 over 2x2x2).
`--config c1|c2|c3|c5` times the other BASELINE configurations on one GPU (C1 HD sod 1-D, C2 2-D Orszag-Tang
1024^2 HLLD + UCT-HLLD, C3 linear waves 128^3 HLLD + PPM + RK3, C5 spherical 384x192x384 Roe + LimO3).
A "step" is one full RK cycle (all stages) = one cell update of every interior cell (src/main.cpp:237-238).

`value`  : device-resident cycles timed with CUDA events on the launching stream, max over ranks.
`e2e`    : the same cycle through the public host API with HOST buffers: pinned-host -> device copy of the
           state (Vc, Vs) a set-up hands over, one host-synchronous cycle, device -> host read of the evolved
           state (Vc, Vs) and of (t, dt, status), every step, all inside the timed region.
           `e2e_upload_only` (extra): the same without the state read-back (result = t, dt, status).
`parity` : before the timed region the same ranks run a SHORT case of the same physics (block-decomposed like
           the bench when N > 1) and compare it with the dump the reference's own CPU build wrote
           (tests/golden/short): max |difference| must be 0.0, else the run fails.
`roofline`: the dominant kernel (the slowest directional sweep), algorithmic bytes
           / CUDA-event duration measured live by idfx_timing, against MEASURED_PEAKS.json.
`cpu_baseline`: the reference's own Kokkos-OpenMP build (oracle/_ref, kind "reference") timed on
           this box's host cores on a bounded sample (a smaller block, a few cycles).
`--impl reference` runs only that CPU arm and prints it in the same JSON shape.
"""
import argparse
import ctypes as C
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DECOMP = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
METRIC = "cell-updates/s (FP64, 3D MHD HLLD+PLM+CT Orszag-Tang, RK2)"


def _grid(*dims):
    return {("Grid", f"X{d + 1}-grid"): list(v) for d, v in enumerate(dims)}


def _u(n, lo="0.0", hi="1.0"):
    return ["1", lo, str(n), "u", hi]


# BASELINE.json configs.  grid(block, nproc) -> ini overrides; `ref` = the reference binary of oracle/_ref that runs
# the same physics on the CPU; `short` = the parity case of tests/cases.py with the same physics.
CONFIGS = {
    "c1": dict(setup="sod", order=2, ref="sod_plm", nvar=3, ndim=1, mhd=False, nstages=2, default_block=500,
               cpu_block=500, short="sod_hllc_plm",
               metric="cell-updates/s (FP64, 1D HD HLLC+PLM Sod tube, RK2)",
               workload="test/HD/sod (1-D Sod shock tube), HLLC + PLM, RK2, outflow",
               grid=lambda b, p: {**_grid(_u(b)), ("Hydro", "solver"): "hllc", ("TimeIntegrator", "tstop"): "1000.0"}),
    "c2": dict(setup="ot2d", order=2, ref="ot2d_plm", nvar=6, ndim=2, mhd=True, nstages=2, default_block=1024,
               cpu_block=1024, short="ot2d_hlld_ucthlld_plm", emf="uct_hlld",
               metric="cell-updates/s (FP64, 2D MHD HLLD+PLM+UCT-HLLD Orszag-Tang, RK2)",
               workload="test/MHD/OrszagTang (2-D Orszag-Tang), HLLD + PLM (vanLeer) + CT(uct_hlld), RK2, periodic",
               grid=lambda b, p: {**_grid(_u(b), _u(b)), ("Hydro", "emf"): "uct_hlld",
                                  ("TimeIntegrator", "tstop"): "1000.0"}),
    "c3": dict(setup="linearwave", order=4, ref="lw_ppm", nvar=8, ndim=3, mhd=True, nstages=3, default_block=128,
               cpu_block=128, short="lw_fast_hlld_ppm_rk3", emf=None,
               metric="cell-updates/s (FP64, 3D MHD HLLD+PPM+CT linear waves, RK3)",
               workload="test/MHD/LinearWaveTest (3-D fast wave), HLLD + PPM + CT(uct_contact), RK3, periodic",
               grid=lambda b, p: {**_grid(_u(b, hi="3.0"), _u(b, hi="1.5"), _u(b, hi="1.5")), ("Hydro", "solver"): "hlld",
                                  ("TimeIntegrator", "nstages"): 3, ("TimeIntegrator", "tstop"): "1000.0"}),
    "c4": dict(setup="ot3d", order=2, ref="ot3d_plm", nvar=8, ndim=3, mhd=True, nstages=2, default_block=512,
               cpu_block=256, short="ot3d_hlld_contact_plm", emf=None, metric=METRIC,
               workload=None,   # formatted with the EMF scheme below
               grid=lambda b, p: {**_grid(_u(b * p[0]), _u(b * p[1]), _u(b * p[2])), ("TimeIntegrator", "tstop"): "1000.0"}),
    "c5": dict(setup="spherical", order=3, ref="sph_limo3", nvar=8, ndim=3, mhd=True, nstages=2, default_block=384,
               cpu_block=96, short=None, emf=None,
               metric="cell-updates/s (FP64, 3D spherical MHD Roe+LimO3+CT, RK2)",
               workload="spherical 3-D MHD (log-r grid, geometric source terms), Roe + LimO3 + CT(uct_contact), RK2",
               grid=lambda b, p: {**_grid(["1", "1.0", str(b), "l", "10.0"],
                                          ["1", "0.7853981633974483", str(b // 2), "u", "2.356194490192345"],
                                          ["1", "0.0", str(b), "u", "6.283185307179586"]),
                                  ("TimeIntegrator", "tstop"): "1000.0"}),
}


def config_of(args):
    cfg = dict(CONFIGS[args.config])
    if args.config == "c4":
        cfg["emf"] = args.emf
        cfg["workload"] = ("test/MHD/OrszagTang3D (3-D Orszag-Tang), HLLD + PLM + CT(%s), RK2, periodic"
                           % (args.emf or "uct_contact"))
    return cfg


def overrides_of(cfg, args, block, nproc):
    o = cfg["grid"](block, nproc)
    if args.config == "c4" and args.emf:
        o[("Hydro", "emf")] = args.emf
    return o


def cells_of(cfg, block, nproc=(1, 1, 1)):
    if cfg["ndim"] == 1:
        return block
    if cfg["ndim"] == 2:
        return block * block
    if cfg["setup"] == "spherical":
        return block * (block // 2) * block
    return block ** 3 * nproc[0] * nproc[1] * nproc[2]


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation on the host cores (oracle/_ref)
# ------------------------------------------------------------------------------------------------
def run_cpu_reference(cfg, args, block, steps, warmup):
    """-> (cell-updates/s over the last `steps` cycles, threads, sample description)"""
    from idefix_b200.input import Input
    from idefix_b200.setups import REGISTRY
    from oracle.ref_run import cleanup, have, run_reference
    if not have(cfg["ref"]):
        raise RuntimeError(f"oracle/_ref/{cfg['ref']}/idefix is missing (build it with oracle/ref_build/build_all.sh)")
    if cfg["setup"] == "spherical":   # the problem directory's own ini (the reference binary was built from it)
        inp = Input(os.path.join(ROOT, "idefix_b200", "host", "problems", "SphericalMHD", "idefix.ini"))
        for name in ("dmp", "vtk"):
            inp.remove("Output", name)
    else:
        inp = Input.from_string(REGISTRY[cfg["setup"]].ini)
    for (blk, name), vals in overrides_of(cfg, args, block, (1, 1, 1)).items():
        inp.set(blk, name, *(vals if isinstance(vals, (list, tuple)) else [vals]))
    inp.set("Output", "log", "1")
    res = run_reference(cfg["ref"], inp.dumps(), maxcycles=warmup + steps + 1, nowrite=True, timeout=3000)
    rates = []
    for line in res["log"].splitlines():
        m = re.match(r"TimeIntegrator:\s+\S+\s+\|\s+(\d+)\s+\|\s+\S+\s+\|\s+([0-9.eE+-]+)", line)
        if m and int(m.group(1)) > warmup:
            rates.append(float(m.group(2)))
    cleanup(res)
    rates = rates[:steps]
    if not rates:
        raise RuntimeError("could not parse the reference log:\n" + res["log"][-2000:])
    value = len(rates) / sum(1.0 / r for r in rates)   # cells*cycles / total time
    sample = (f"{cfg['workload']}, {cells_of(cfg, block)} cells (block {block}), {len(rates)} cycles after {warmup} "
              f"warm-up, reference Kokkos-OpenMP build ({cfg['ref']}), {res['threads']} threads")
    return value, res["threads"], sample


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = config_of(args)
    block = args.cpu_block or cfg["cpu_block"]
    t0 = time.time()
    value, threads, sample = run_cpu_reference(cfg, args, block, args.steps, args.warmup)
    cells = cells_of(cfg, block)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cells / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # what this arm actually ran: ONE OpenMP process on the host cores, a bounded block of the same workload
        "config": {"workload": cfg["workload"], "block": block, "cells": cells,
                   "sample": "one Kokkos-OpenMP process on the host cores, no decomposition, no halo exchange; "
                             "bounded sample of the GPU arm's workload (same physics, smaller block)"},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(cfg, args, block, nproc):
    out = {"workload": cfg["workload"], "baseline_config": args.config.upper()}
    if cfg["ndim"] == 3 and cfg["setup"] != "spherical":
        out.update({"block_per_gpu": [block] * 3, "global_grid": [block * nproc[0], block * nproc[1], block * nproc[2]],
                    "decomposition": list(nproc)})
    else:
        out.update({"block": block, "cells": cells_of(cfg, block)})
    out["halo"] = "none" if args.gpus == 1 else args.halo
    out["l2"] = ("inputs (>= 50 GB of state arrays at 512^3) far larger than the 126 MB L2; no flush needed"
                 if cells_of(cfg, block) * cfg["nvar"] * 8 * 4 > 4e8 else
                 "state arrays smaller than the 126 MB L2: an L2-resident workload by construction (no flush: the "
                 "reference's own benchmark of this configuration runs the same way)")
    return out


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._th = index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._loop, daemon=True)
        self._th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for name, flag in zip(names, s[2:6]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def sweep_algorithmic_bytes(nvar, ndim, mhd, emf, stage0_share, direction):
    """DESIGN.md 'algorithmic bytes': per interior cell and per sweep launch (average over the RK stages).
    Every sweep reads Vc (nvar) (+ the normal Vs and the face EMF outputs in MHD: 2 EMFs + contact weight, or
    2 + 4 UCT coefficients).  The x1 sweep rebuilds Uc from Vc: it WRITES the non-CT conservative variables
    (and their copy Uc0 on stage 0), InvDt (stage 0) and, for uct_contact, the cell-centred EMFs; the x2/x3
    sweeps read-modify-write Uc and InvDt (stage 0)."""
    face = {"uct_contact": 3, None: 3, "uct0": 2, "arithmetic": 2, "uct_hll": 6, "uct_hlld": 6}[emf] if mhd else 0
    if mhd and ndim == 1:
        face = 0
    nu = nvar - (ndim if mhd else 0)
    normal = 1 if mhd else 0
    if direction == 0:
        centre = (3 if ndim == 3 else 1) if (mhd and ndim >= 2 and emf in (None, "uct_contact")) else 0
        doubles = nvar + normal + nu + nu * stage0_share + face + centre + 1 * stage0_share
    else:
        doubles = nvar + normal + 2 * nu + face + 2 * stage0_share
    return 8.0 * doubles


def short_case_parity(cfg, nproc, xproc, rank, world, local):
    """Runs the SHORT parity case of the config's physics on the bench's own ranks (decomposed like the bench) and
    returns max |difference| against the dump of the reference's CPU build (tests/golden/short), on rank 0."""
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    from cases import SHORT
    from idefix_b200.dump import readDump
    from idefix_b200.setups import REGISTRY
    case = [c for c in SHORT if c.name == cfg["short"]][0]
    s = REGISTRY[case.setup](order=case.order, overrides=case.overrides, nproc=nproc, xproc=xproc, device=local)
    ex = None
    if world > 1:
        from idefix_b200.halo import HaloExchanger
        ex = HaloExchanger(s.data, nproc, xproc)
    s.initialise()
    ncyc = max(case.cycles)
    parity.advance_cycles(s, ncyc, True)
    mine = {k: np.ascontiguousarray(v) for k, v in parity.fields_of(s).items()}
    mathexc = s.data.math_exceptions()
    npi = [s.grid.np_int[a] // nproc[a] for a in range(3)]
    gathered = [(xproc, mine)]
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((xproc, mine), gathered, dst=0)
    worst = None
    if rank == 0:
        gold = readDump(os.path.join(ROOT, "tests", "golden", "short", f"{case.name}.c{ncyc:04d}.dmp"))
        worst = 0.0
        for fname, ref in gold.data.items():
            if not fname.startswith(("Vc-", "Vs-")):
                continue
            full = np.full(np.asarray(ref).shape, np.nan)
            for xp, fields in gathered:
                blk = fields[fname]
                lo = [xp[a] * npi[a] for a in range(3)]
                full[lo[0]:lo[0] + blk.shape[0], lo[1]:lo[1] + blk.shape[1], lo[2]:lo[2] + blk.shape[2]] = blk
            diff = np.abs(full - np.asarray(ref))
            worst = max(worst, float(np.nan_to_num(diff, nan=np.inf).max()))
    if ex is not None:
        ex.close()
    s.data.close()
    return {"case": case.name, "cycles": ncyc, "decomposition": list(nproc), "max_abs": worst,
            "math_exceptions": int(mathexc), "against": "tests/golden/short (dumps of the reference's CPU build)"}


def timed_cycles(lib, h, stream, steps, barrier, world):
    import torch
    import torch.distributed as dist
    from idefix_b200 import _abi
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    _abi.check(lib.idfx_cycles(h, steps))
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main_gpu(args):
    import torch
    import torch.distributed as dist

    from idefix_b200 import _abi
    from idefix_b200.setups import REGISTRY

    cfg = config_of(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
        args.gpus = world
    if world > 1 and args.config != "c4":
        raise SystemExit("bench.py: only --config c4 is block-decomposed over several GPUs")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nproc = DECOMP[world]
    xproc = (rank % nproc[0], (rank // nproc[0]) % nproc[1], rank // (nproc[0] * nproc[1]))
    b = args.block or cfg["default_block"]
    lib = _abi.lib()
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- parity gate: a short case of the same physics on these very ranks against the reference's dump ----------
    parity_res = None
    if cfg["short"] and not args.no_parity:
        parity_res = short_case_parity(cfg, nproc, xproc, rank, world, local)
        if rank == 0 and not (parity_res["max_abs"] == 0.0 and parity_res["math_exceptions"] == 0):
            raise SystemExit(f"bench.py: parity gate failed before the timed region: {parity_res}")

    def make_block(block):
        setup = REGISTRY[cfg["setup"]](order=cfg["order"], overrides=overrides_of(cfg, args, block, nproc),
                                       nproc=nproc, xproc=xproc, device=local)
        _abi.check(lib.idfx_set_stream(setup.data._handle, C.c_void_p(stream.cuda_stream)))
        ex = None
        if world > 1:
            from idefix_b200.halo import HaloExchanger
            ex = HaloExchanger(setup.data, nproc, xproc, mode=args.halo)
        setup.InitFlow(setup.data)
        barrier()   # the first exchange should not wait for a rank still filling its initial condition
        setup.data.SetBoundaries()
        return setup, ex

    setup, exchanger = make_block(b)
    data = setup.data
    h = data._handle

    # ---- warm-up --------------------------------------------------------------------------------
    _abi.check(lib.idfx_cycles(h, args.warmup))
    barrier()

    # ---- timed region: K device-resident cycles ----------------------------------------------------
    _abi.check(lib.idfx_timing(h, 1, None))
    l1 = data.launch_count()
    with ClockSampler(local) as clocks:
        ms = timed_cycles(lib, h, stream, args.steps, barrier, world)
    launches = data.launch_count() - l1
    tim = (C.c_double * 12)()
    _abi.check(lib.idfx_timing(h, 2, tim))
    _abi.check(lib.idfx_timing(h, 0, None))
    cells_global = cells_of(cfg, b, nproc)
    value = cells_global * args.steps / (ms * 1e-3)

    # sanity: the run must have produced finite data and a solenoidal field
    nan = data.CheckNan()
    divb = data.hydro.CheckDivB() if cfg["mhd"] else 0.0
    mathexc = data.math_exceptions()
    if nan or not divb < 1e-10 or mathexc:
        raise SystemExit(f"bench.py: invalid state after the timed region (nan={nan}, divB={divb}, "
                         f"fast-math range-check failures={mathexc})")

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    classes = ["sweep_x1", "sweep_x2", "sweep_x3", "constrained_transport", "stage_epilogue", "boundaries"]
    kern = {c: {"launch_brackets": int(tim[2 * i]), "ms_total": tim[2 * i + 1]} for i, c in enumerate(classes)}
    sweeps = [(tim[2 * i + 1] / max(tim[2 * i], 1), i) for i in range(cfg["ndim"])]
    dur_ms, top = max(sweeps)
    nstages = cfg["nstages"]
    emf = cfg.get("emf")
    bytes_cell = sweep_algorithmic_bytes(cfg["nvar"], cfg["ndim"], cfg["mhd"], emf, 1.0 / nstages, top)
    alg_bytes = bytes_cell * cells_of(cfg, b)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (dur_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "sweep_dram_bytes.json")
    if args.config == "c4" and os.path.exists(tpath):
        try:
            with open(tpath) as fh:
                tj = json.load(fh)
            if tj.get("block") == b and tj.get("emf") == (emf or "uct_contact"):
                traffic = tj["bytes_per_launch"][classes[top]]
        except Exception:
            traffic = None
    kname = ["pencilRowKernel (x1 sweep: reconstruct + Riemann + EMF store + RHS + prim->cons, fused)",
             "pencilMarchKernel<JDIR> (x2 sweep: reconstruct + Riemann + EMF store + RHS, fused)",
             "pencilMarchKernel<KDIR> (x3 sweep: reconstruct + Riemann + EMF store + RHS, fused)"][top]
    # B_alg of the whole cycle: 8 B x (n_c + n_s) x (2 + 3 (nstages - 1))
    cycle_alg = 8 * cfg["nvar"] * (2 + 3 * (nstages - 1))
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_bytes_per_cell": bytes_cell, "avg_launch_ms": dur_ms,
                "cycle_alg_bytes_per_cell_update": cycle_alg,
                "cycle_frac_of_peak": value / world * cycle_alg / (peak * 1e9),
                "kernel_ms_total": kern, "timed_region_ms": ms}

    # ---- e2e: host buffers in, host buffers out, every step ----------------------------------------------
    fields = [_abi.FIELD_VC] + ([_abi.FIELD_VS] if cfg["mhd"] else [])
    hbuf = []
    for f in fields:
        dims, _ = data.field_dims(f)
        t = torch.empty(dims, dtype=torch.float64, pin_memory=True)
        _abi.check(lib.idfx_download(h, f, C.c_void_p(t.data_ptr())))
        hbuf.append(t)
    state_bytes = sum(t.numel() for t in hbuf) * 8
    tt, dd, ff = C.c_double(), C.c_double(), C.c_int()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def e2e_run(nsteps, read_back):
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(nsteps):
            # the step's inputs: the state a set-up / restart hands over, from pinned host memory
            for f, t in zip(fields, hbuf):
                _abi.check(lib.idfx_upload(h, f, C.c_void_p(t.data_ptr())))
            # one cycle through the host-synchronous entry point; (t, dt, status flags) -- what the reference's main
            # loop looks at after TimeIntegrator::Cycle -- come back from the device ...
            _abi.check(lib.idfx_cycle(h, C.byref(tt), C.byref(dd), C.byref(ff)))
            if read_back:   # ... and so does the evolved state (DataBlockHost::SyncFromDevice)
                for f, t in zip(fields, hbuf):
                    _abi.check(lib.idfx_download(h, f, C.c_void_p(t.data_ptr())))
        ev1.record(stream)
        barrier()
        e_ms = max(ev0.elapsed_time(ev1), (time.perf_counter() - t0) * 1e3)
        if world > 1:
            t = torch.tensor([e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        return e_ms

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    up_steps = max(1, min(3, e2e_steps))
    up_ms = e2e_run(up_steps, False)
    e2e_ms = e2e_run(e2e_steps, not args.e2e_upload_only)
    full = not args.e2e_upload_only
    e2e = {"value": cells_global * e2e_steps / (e2e_ms * 1e-3), "unit": "cell-updates/s",
           "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": (state_bytes if full else 0) + 20,
           "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
           "api": "DataBlockHost.SyncToDevice (idfx_upload Vc,Vs from pinned host memory) -> idfx_cycle -> (t, dt, status)"
                  + (" + DataBlockHost.SyncFromDevice (idfx_download Vc,Vs into pinned host memory)" if full else "")}
    e2e_upload_only = {"value": cells_global * up_steps / (up_ms * 1e-3), "unit": "cell-updates/s",
                       "h2d_bytes_per_step": state_bytes, "d2h_bytes_per_step": 20, "steps": up_steps,
                       "ms_per_step": up_ms / up_steps}

    # ---- N > 1: the same measurement at 256^3 per GPU (at N = 8: BASELINE config C4 as stated, 512^3 over 2x2x2) ----
    stated = None
    if world > 1 and args.config == "c4" and b != 256 and not args.no_stated:
        if exchanger is not None:
            exchanger.close()
        data.close()
        data = None
        setup2, ex2 = make_block(256)
        h2 = setup2.data._handle
        _abi.check(lib.idfx_cycles(h2, args.warmup))
        ms2 = timed_cycles(lib, h2, stream, args.steps, barrier, world)
        cells2 = cells_of(cfg, 256, nproc)
        stated = {"block_per_gpu": [256] * 3, "global_grid": [256 * nproc[0], 256 * nproc[1], 256 * nproc[2]],
                  "decomposition": list(nproc), "value": cells2 * args.steps / (ms2 * 1e-3), "unit": "cell-updates/s",
                  "ms_per_step": ms2 / args.steps, "steps": args.steps,
                  "divB": setup2.data.hydro.CheckDivB(), "nan": setup2.data.CheckNan()}
        if ex2 is not None:
            ex2.close()
        setup2.data.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only) --------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            v, threads, sample = run_cpu_reference(cfg, args, args.cpu_block or cfg["cpu_block"], args.cpu_steps, 1)
            cpu = {"value": v, "unit": "cell-updates/s", "cores": threads, "kind": "reference", "sample": sample}
        except Exception as e:  # the CPU arm is a reported baseline; its absence must be visible, not fatal
            cpu = {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference",
                   "sample": f"unavailable: {e}"[:300]}

    if rank == 0:
        line = {
            "metric": cfg["metric"], "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, args, b, nproc), "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "e2e_upload_only": e2e_upload_only, "parity": parity_res,
            "parity_ngpu_max_abs": parity_res["max_abs"] if parity_res else None,
            "per_gpu": value / world, "divB": divb,
            "published_h100_per_gpu": 3.08e8,
        }
        if stated is not None:
            line["c4_as_stated"] = stated
        print(json.dumps(line), flush=True)
    if data is not None:
        if exchanger is not None:
            exchanger.close()
        data.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS), help="BASELINE.json configuration (default C4)")
    ap.add_argument("--block", type=int, default=0, help="interior cells per direction per GPU (default: the config's)")
    ap.add_argument("--emf", default=None, help="C4: EMF averaging ([Hydro] emf); default = the reference's default uct_contact")
    ap.add_argument("--halo", default="p2p", choices=["nccl", "p2p"],
                    help="N > 1: p2p = the library's own exchange kernels over NVLink peer memory (default), "
                         "nccl = packed slabs through torch.distributed point-to-point")
    ap.add_argument("--e2e-steps", type=int, default=1000000, help="cap on the e2e steps (default: K)")
    ap.add_argument("--e2e-upload-only", action="store_true",
                    help="e2e without the read-back of the evolved state (default: full state up AND down every step)")
    ap.add_argument("--e2e-full-state", action="store_true", help="(default behaviour; kept for older scripts)")
    ap.add_argument("--cpu-block", type=int, default=0, help="block of the bounded CPU sample (default: the config's)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity gate (profiling runs under ncu only)")
    ap.add_argument("--no-stated", action="store_true", help="N > 1: skip the additional 256^3-per-GPU measurement")
    args = ap.parse_args()
    if args.gpus not in DECOMP:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
